// explicit instantiations of the fused integrator: double, 3 gas(es), all alpha modes, and the
// specialised per-gas forms of the default (EXP) mode
#include "ufair_kernel.cuh"
namespace ufair {
UFAIR_DEFINE_LAUNCH_EXP(double, 3, UFAIR_TRY_FORM(double, 3, kForms3[0]))
UFAIR_DEFINE_LAUNCH(double, 3, UFAIR_ALPHA_SINH)
UFAIR_DEFINE_LAUNCH(double, 3, UFAIR_ALPHA_NEWTON)
UFAIR_DEFINE_LAUNCH(double, 3, UFAIR_ALPHA_ONE)
}  // namespace ufair
