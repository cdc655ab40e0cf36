// Device-side barrier between the GPUs of one node over NVLink peer memory (view-sharded mode).
//
// Every participant owns a small flag array in SYMMETRIC memory (same allocation on each GPU, all peers' copies mapped into
// this GPU's address space).  mdb_peer_barrier launches one warp: it bumps a local epoch counter, writes the epoch into its
// slot of every peer's flag array (st.release.sys through NVLink), then spins until every peer's slot in its OWN array has
// reached the epoch (ld.acquire.sys).  Stream order + the system-scope release/acquire make everything the peers wrote
// before their barrier call (the K/V projection of a multiview block) visible to the kernels this GPU launches after it —
// which read that K/V in place through TMA loads on the peer-mapped addresses (mdb_attention_multi).  The kernel is a plain
// stream operation, so it is captured into the denoising step's CUDA graph; the epoch lives in device memory and advances
// on every replay.  No NCCL call is on the data path.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/magicdrive_b200.h"
#include "common_host.h"

namespace {

__global__ void peer_barrier_kernel(uint32_t* const* flag_ptrs, int rank, int world, int channel, int n_channels, uint32_t* epoch,
                                    long long timeout_cycles, int* timed_out) {
  const int lane = threadIdx.x;
  uint32_t e = 0;
  if (lane == 0) {
    e = epoch[channel] + 1;
    epoch[channel] = e;
  }
  e = __shfl_sync(0xffffffffu, e, 0);
  __threadfence_system();  // this GPU's earlier writes (previous kernels of the stream) before the flag becomes visible
  if (lane < world && lane != rank) {
    uint32_t* dst = flag_ptrs[lane] + (channel * world + rank);
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(e) : "memory");
  }
  if (lane < world && lane != rank) {
    const uint32_t* src = flag_ptrs[rank] + (channel * world + lane);
    const long long t0 = clock64();
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
      if (timeout_cycles > 0 && clock64() - t0 > timeout_cycles) {
        *timed_out = 1;  // a peer never arrived: report instead of hanging the GPU (the host checks the flag)
        break;
      }
    } while (static_cast<int32_t>(v - e) < 0);
  }
  __syncwarp();
  __threadfence_system();
}

}  // namespace

extern "C" int mdb_peer_barrier(void* const* flag_ptrs_dev, int rank, int world, int channel, int n_channels, void* epoch_dev,
                                long long timeout_cycles, int* timed_out_dev, void* stream) {
  using namespace mdb;
  if (!flag_ptrs_dev || !epoch_dev || !timed_out_dev) return set_error(MDB_ERR_INVALID, "mdb_peer_barrier: null pointer");
  if (world < 1 || world > 32 || rank < 0 || rank >= world || channel < 0 || channel >= n_channels)
    return set_error(MDB_ERR_INVALID, "mdb_peer_barrier: bad rank/world/channel (%d/%d/%d)", rank, world, channel);
  peer_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<uint32_t* const*>(flag_ptrs_dev), rank, world,
                                                                       channel, n_channels, static_cast<uint32_t*>(epoch_dev),
                                                                       timeout_cycles, timed_out_dev);
  MDB_CHECK_LAUNCH("peer_barrier_kernel");
  return MDB_OK;
}
