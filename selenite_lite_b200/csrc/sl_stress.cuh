// Stress build (-DSL_TC_STRESS; lib/libselenite_b200_stress.so, tests/test_gpu_stress_build.py): every hand-over point of the
// persistent kernels (mbarrier arrive / wait, named barrier, look-back poll) first sleeps a pseudo-random time of up to ~16 us in about
// one call out of four, per warp — roles and warps of one role drift apart by several tiles' worth of time, which is what it takes to
// show a hand-over protocol that only holds for the usual timing (round 2: three such faults in the q15 tensor-core kernel). The
// GPU tests must pass unchanged on that build. In the normal build sl_jitter () is empty.
#pragma once
#ifdef SL_TC_STRESS
__device__ __forceinline__ void sl_jitter ()
{
  unsigned t; asm volatile ("mov.u32 %0, %%clock;" : "=r"(t));
  const unsigned h = (t ^ ((threadIdx.x >> 5) * 2654435761u)) * 2246822519u;
  if ((h >> 28) < 4u) __nanosleep ((h >> 8) & 0x3FFFu);
}
#else
__device__ __forceinline__ void sl_jitter () {}
#endif
