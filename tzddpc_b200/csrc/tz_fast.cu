// fast_step_kernel (tz_fast.cuh) in its own translation unit: it compiles in seconds, step_kernel<B0> in minutes.
#include "tz_fast.cuh"

namespace tz {
template int launch_fast<B0>(const TzProgram*, const SolverParams&, const StepArgs&, cudaStream_t);
}
