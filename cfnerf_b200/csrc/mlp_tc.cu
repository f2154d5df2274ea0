// placeholder until the tcgen05 kernel lands (replaced in the next commit)
#include "handle.h"
namespace cfn {
int tc_create(CfnHandle*) { set_error("tensor-core path not built"); return CFN_EINVAL; }
void tc_destroy(CfnHandle*) {}
int tc_pack(CfnHandle*, cudaStream_t) { return CFN_EINVAL; }
size_t tc_workspace_bytes(const CfnHandle*, int64_t) { return 0; }
int tc_network_fwd(CfnHandle*, const float*, const float*, const float*, const float*, int64_t, int, float*, void*, size_t, cudaStream_t) { return CFN_EINVAL; }
}
