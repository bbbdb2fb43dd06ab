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
//
// Two ways to publish, chosen by batch size (eb_control_dev_gather):
//  * fused: the solve kernel stores its rows itself (above).  With many waves per SM the
//    stores and fences of finished warps overlap the math of the others (+2.5 % at 65 536
//    instances per GPU).
//  * side stream: a batch that is a single wave (4096 instances: a 31 us kernel) has nothing
//    to overlap with -- the NVLink round trips of the publication would sit at the kernel's
//    tail (+34 us at 8 GPUs).  There the solve kernel writes u0 locally and
//    peer_publish_kernel, on the group's own stream, copies the block to every rank and raises
//    the flags while the next step's solve kernel already runs.
#pragma once

#include <cuda_runtime.h>

namespace eb
{
constexpr int kPublishPeers = 8;

struct PublishParams
{
  const double* src;                          // this rank's first twists of the step, [elems]
  long long elems;
  int n_peer;
  double* dst[kPublishPeers];                 // every rank's gathered buffer, offset to this rank's block
  unsigned long long* flag[kPublishPeers];    // this rank's slot in every rank's arrival flags
  unsigned long long flag_value;
  const unsigned long long* my_flags;         // reuse guard (see solve_kernel.cuh)
  unsigned long long need;
  unsigned int* done_counter;
};

__global__ void __launch_bounds__(256) peer_publish_kernel(const PublishParams p)
{
  if (threadIdx.x < p.n_peer)
    while (*(const volatile unsigned long long*)(p.my_flags + threadIdx.x) < p.need) __nanosleep(100);
  __syncthreads();
  const long long pairs = p.elems >> 1;  // elems = 3 * batch; 16-byte pieces, a possible odd tail below
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += (long long)gridDim.x * blockDim.x)
  {
    const double2 v = reinterpret_cast<const double2*>(p.src)[i];
    for (int q = 0; q < p.n_peer; q++) reinterpret_cast<double2*>(p.dst[q])[i] = v;
  }
  if ((p.elems & 1) && blockIdx.x == 0 && threadIdx.x == 0)
    for (int q = 0; q < p.n_peer; q++) p.dst[q][p.elems - 1] = p.src[p.elems - 1];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0)
  {
    const unsigned int prev = atomicAdd(p.done_counter, 1u);
    if (prev == gridDim.x - 1)
    {
      *p.done_counter = 0u;
      __threadfence_system();
      for (int q = 0; q < p.n_peer; q++) *(volatile unsigned long long*)p.flag[q] = p.flag_value;
    }
  }
}

// Synchronous form for single-wave batches: publish AND wait in one kernel on the controller's own stream, right
// behind the solve kernel -- no side-stream hop, no second launch.  Nothing else needs the SMs between the solve
// kernel and the consumer of the gathered rows, so the copy uses many blocks; the last block to finish its stores
// fences, raises this rank's flags on every peer and then waits for the peers' flags.
__global__ void __launch_bounds__(256) peer_publish_wait_kernel(const PublishParams p)
{
  // programmatic dependent launch: the solve kernel ahead has completed (its rows are in p.src) past this point
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x < p.n_peer)
    while (*(const volatile unsigned long long*)(p.my_flags + threadIdx.x) < p.need) __nanosleep(100);
  __syncthreads();
  const long long pairs = p.elems >> 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += (long long)gridDim.x * blockDim.x)
  {
    const double2 v = reinterpret_cast<const double2*>(p.src)[i];
    for (int q = 0; q < p.n_peer; q++) reinterpret_cast<double2*>(p.dst[q])[i] = v;
  }
  if ((p.elems & 1) && blockIdx.x == 0 && threadIdx.x == 0)
    for (int q = 0; q < p.n_peer; q++) p.dst[q][p.elems - 1] = p.src[p.elems - 1];
  __threadfence();
  __syncthreads();
  __shared__ unsigned int s_last;
  if (threadIdx.x == 0)
  {
    const unsigned int prev = atomicAdd(p.done_counter, 1u);
    s_last = prev == gridDim.x - 1 ? 1u : 0u;
    if (s_last)
    {
      *p.done_counter = 0u;
      __threadfence_system();
      for (int q = 0; q < p.n_peer; q++) *(volatile unsigned long long*)p.flag[q] = p.flag_value;
    }
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x < p.n_peer)
    while (*(const volatile unsigned long long*)(p.my_flags + threadIdx.x) < p.flag_value) __nanosleep(64);
  __syncthreads();
  __threadfence_system();  // acquire: the peers' rows are visible to what follows on this stream
}

// spins (one thread) until every rank's arrival flag has reached `step`
__global__ void peer_wait_kernel(const volatile unsigned long long* flags, int world, unsigned long long step)
{
  for (int r = threadIdx.x; r < world; r += blockDim.x)
    while (flags[r] < step) __nanosleep(200);
  __syncthreads();
  __threadfence_system();  // acquire: the rows published before the flags are visible to what follows
}
}  // namespace eb
