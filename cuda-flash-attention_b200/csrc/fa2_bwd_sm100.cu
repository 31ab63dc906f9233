// placeholder until the tcgen05 backward lands
#include "fa2_common.h"
namespace fa2 {
cudaError_t launch_bwd(const BwdParams&, cudaStream_t) { return cudaErrorNotSupported; }
}
