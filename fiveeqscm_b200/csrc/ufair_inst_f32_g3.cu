// explicit instantiations of the fused integrator: float, 3 gas(es), all alpha modes, and the
// specialised per-gas forms of the default (EXP) mode
#include "ufair_kernel.cuh"
namespace ufair {
UFAIR_DEFINE_LAUNCH_EXP(float, 3, UFAIR_TRY_FORM(float, 3, kForms3[0]))
UFAIR_DEFINE_LAUNCH(float, 3, UFAIR_ALPHA_SINH)
UFAIR_DEFINE_LAUNCH(float, 3, UFAIR_ALPHA_NEWTON)
UFAIR_DEFINE_LAUNCH(float, 3, UFAIR_ALPHA_ONE)
}  // namespace ufair
