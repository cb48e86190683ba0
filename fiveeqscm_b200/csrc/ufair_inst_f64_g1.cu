// explicit instantiations of the fused integrator: double, 1 gas(es), all alpha modes
#include "ufair_kernel.cuh"
namespace ufair {
UFAIR_DEFINE_LAUNCH(double, 1, UFAIR_ALPHA_EXP)
UFAIR_DEFINE_LAUNCH(double, 1, UFAIR_ALPHA_SINH)
UFAIR_DEFINE_LAUNCH(double, 1, UFAIR_ALPHA_NEWTON)
UFAIR_DEFINE_LAUNCH(double, 1, UFAIR_ALPHA_ONE)
}  // namespace ufair
