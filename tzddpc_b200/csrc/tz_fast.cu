// fast_step_kernel (tz_fast.cuh) in its own translation unit: it compiles in seconds, step_kernel<B0> in minutes.
#include "tz_fast.cuh"

namespace tz {
template int launch_fast<B0>(const TzProgram*, const SolverParams&, const StepArgs&, cudaStream_t);
template int launch_fast_set<B0>(const TzProgram*, const SolverParams&, const StepArgs&, const SetEntry*, const int32_t*, int, cudaStream_t);
}
