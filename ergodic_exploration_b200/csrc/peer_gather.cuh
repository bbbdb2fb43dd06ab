// peer_gather.cuh -- the multi-GPU result exchange of the sharded control() step,
// fused into the solve kernel over NVLink / NVSwitch peer memory (sm_100a).
//
// the instances are block-partitioned over the ranks and the
// only exchange is the gather of the first twists u0 (B/g x 3 doubles per rank and
// step).  A collective call per step costs more than the 30 us kernel it follows, so
// there is no collective: every rank maps every other rank's gathered buffer
// (cudaIpc handles exchanged once at start-up) and the solve kernel stores each
// instance's u0 directly into all of them (solve_kernel.cuh, "fused gather").  Arrival
// is tracked by one 64-bit flag per (receiver, sender): the sender's last warp raises
// it to the step number after a system-scope fence.  A consumer that needs step s on
// rank r enqueues peer_wait_kernel, which spins until all of r's flags are >= s.
// kPeerBuffers gathered buffers rotate by step; before a warp stores into the buffer last
// used kPeerBuffers steps ago it checks (its own, local flags) that every rank has finished
// the step after that one -- a rank reads a step's rows before launching its next step, so
// nobody can still be reading.  In a balanced run that condition is long true: the steps
// of different ranks are not coupled and no wait kernel sits between launches.
#pragma once

#include <cuda_runtime.h>

namespace eb
{
// spins (one thread) until every rank's arrival flag has reached `step`
__global__ void peer_wait_kernel(const volatile unsigned long long* flags, int world, unsigned long long step)
{
  for (int r = threadIdx.x; r < world; r += blockDim.x)
    while (flags[r] < step) __nanosleep(200);
  __syncthreads();
  __threadfence_system();  // acquire: the rows published before the flags are visible to what follows
}
}  // namespace eb
