// dist_engine.cu -- tensor-core (tcgen05) distance engine shared by KNeighbors and SVC.  Placeholder until
// the engine lands: every handle reports "not usable" and the fp64 CUDA-core kernels run.
#include "common.h"

namespace tcsdn {

int engine_create(tcsdn_model *m) { m->engine = nullptr; return TCSDN_OK; }
void engine_destroy(tcsdn_model *m) { m->engine = nullptr; }
bool engine_usable(const tcsdn_model *, int64_t) { return false; }
int launch_engine(tcsdn_model *, const void *, int64_t, int, int32_t *, double *, cudaStream_t) {
    set_error("tensor-core engine not built");
    return TCSDN_EINVAL;
}

}  // namespace tcsdn
