// explicit instantiations of the fused integrator: float, 2 gas(es), all alpha modes
#include "ufair_kernel.cuh"
namespace ufair {
UFAIR_DEFINE_LAUNCH(float, 2, UFAIR_ALPHA_EXP)
UFAIR_DEFINE_LAUNCH(float, 2, UFAIR_ALPHA_SINH)
UFAIR_DEFINE_LAUNCH(float, 2, UFAIR_ALPHA_NEWTON)
UFAIR_DEFINE_LAUNCH(float, 2, UFAIR_ALPHA_ONE)
}  // namespace ufair
