// search.cu — placeholder until the search driver lands (next commit): fails loudly.
#include "vsg_internal.h"
using namespace vsg;
extern "C" void vsg_search_opts_default(vsg_search_opts * o)
{
  o->id = 0.0; o->weak_id = 10.0; o->maxaccepts = 1; o->maxrejects = 32; o->wordlength = 8;
  o->minwordmatches = -1; o->iddef = 2; o->strand_both = 0; o->mask_lower = 0; o->reserved = 0;
}
extern "C" int vsg_search_batch(vsg_ctx *, const vsg_index *, const vsg_seqset *, const vsg_seqset *, int64_t,
                                int64_t, const vsg_search_opts *, vsg_search_result *, int, int32_t *, int64_t *)
{ Error::set("vsg_search_batch: not implemented yet"); return VSG_EINVAL; }
