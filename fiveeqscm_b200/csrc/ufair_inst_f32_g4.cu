// explicit instantiations of the fused integrator: float, 4 gas(es), all alpha modes, and the
// specialised per-gas forms of the default (EXP) mode
#include "ufair_kernel.cuh"
namespace ufair {
UFAIR_DEFINE_LAUNCH_EXP(float, 4, UFAIR_TRY_FORM(float, 4, kForms4[0]))
UFAIR_DEFINE_LAUNCH(float, 4, UFAIR_ALPHA_SINH)
UFAIR_DEFINE_LAUNCH(float, 4, UFAIR_ALPHA_NEWTON)
UFAIR_DEFINE_LAUNCH(float, 4, UFAIR_ALPHA_ONE)
}  // namespace ufair
