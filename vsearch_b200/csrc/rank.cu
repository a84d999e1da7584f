// rank.cu — placeholder until the ranker kernels land (next commit): fails loudly.
#include "vsg_internal.h"
using namespace vsg;
struct vsg_index { int dummy; };
extern "C" int vsg_index_create(vsg_ctx *, const vsg_seqset *, int, int, vsg_index ** out)
{ if (out) { *out = nullptr; } Error::set("vsg_index_create: not implemented yet"); return VSG_EINVAL; }
extern "C" void vsg_index_destroy(vsg_index *) {}
extern "C" int vsg_rank(vsg_ctx *, const vsg_index *, const vsg_seqset *, int64_t, int64_t, int, int, int,
                        uint32_t *, uint32_t *, int32_t *)
{ Error::set("vsg_rank: not implemented yet"); return VSG_EINVAL; }
