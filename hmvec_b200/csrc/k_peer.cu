// k_peer.cu -- the one exchange of a z-sharded run (the P(k,z) tables Limber integrates over ALL redshifts,
// cosmology.py:867-904) done by the kernel that forms the tables: every rank sums P = P1h + P2h of its redshift slab and
// stores the result straight into the gathered [nz_total][nsp][nk] table of EVERY rank over NVLink (peer memory mapped
// through CUDA IPC), then raises a per-rank step flag on every peer; a one-CTA kernel in front of the Limber kernels
// waits for the flags.  No collective launch, no staging buffer, no rendezvous on the host: pack + all-gather + the
// ordering are two launches of this file.  NCCL (zshard.ZComm.all_gather_rows) remains the fallback when peer access
// is not available.
//
// Memory ordering: every thread's peer stores are followed by a system-scope fence; the last CTA to finish (device
// counter) publishes `step` to flag[rank] on each peer with a release store; the waiter reads with acquire loads.
// Buffers are used alternately (step parity): a rank can only reach the stores of step s+2 after every peer has
// signalled step s+1, which a peer does after its own consumers of step s were queued on the same stream.
#include "common.cuh"

namespace hmv {

struct PeerArgs {
  const double* a[4];
  const double* b[4];
  double* buf[HMV_MAX_PEERS];                  // gathered table of each rank (this step's parity half)
  unsigned long long* flag[HMV_MAX_PEERS];     // flag array [HMV_MAX_PEERS] of each rank
  int nrow, ncol, nsp, npeers, rank;
  long long row0;                              // first global row of this rank's slab
  unsigned long long step;
  unsigned int* done;                          // local CTA counter (self-resetting)
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// grid (ceil(ncol / 512), nrow), 256 threads: a thread owns two adjacent columns of one row of every spectrum
__global__ void __launch_bounds__(256) peer_scatter_kernel(const PeerArgs q) {
  const int k = 2 * (blockIdx.x * blockDim.x + threadIdx.x), z = blockIdx.y;
  if (k < q.ncol) {
    const long long i = (long long)z * q.ncol + k;
    const bool two = k + 1 < q.ncol;
    const bool vec = two && ((q.ncol & 1) == 0);             // 16-byte stores need even rows (bases are 256-B aligned)
    for (int s = 0; s < q.nsp; ++s) {
      double v0 = q.a[s][i], v1 = two ? q.a[s][i + 1] : 0.0;
      if (q.b[s]) { v0 += q.b[s][i]; if (two) v1 += q.b[s][i + 1]; }
      const long long o = ((q.row0 + z) * q.nsp + s) * (long long)q.ncol + k;
      for (int p = 0; p < q.npeers; ++p) {
        double* dst = q.buf[p] + o;
        if (vec) *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
        else { dst[0] = v0; if (two) dst[1] = v1; }
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int total = gridDim.x * gridDim.y;
    const unsigned int prev = atomicAdd(q.done, 1u);
    if (prev == total - 1u) {
      __threadfence_system();
      *q.done = 0u;                                          // the next launch on this stream starts from zero
      for (int p = 0; p < q.npeers; ++p) st_release_sys(q.flag[p] + q.rank, q.step);
    }
  }
}

// one CTA, thread r waits for rank r's flag; status: 0 ok, 1 + r = rank r never arrived within the timeout
__global__ void peer_wait_kernel(const unsigned long long* flags, int npeers, unsigned long long step,
                                 long long timeout_cycles, int* status) {
  const int r = threadIdx.x;
  if (r >= npeers) return;
  if (*reinterpret_cast<volatile int*>(status) != 0) return;     // an earlier wait gave up: fail fast, the host raises
  const long long t0 = clock64();
  while (ld_acquire_sys(flags + r) < step) {
    if (clock64() - t0 > timeout_cycles) { atomicExch(status, 1 + r); break; }
    __nanosleep(200);
  }
}

}  // namespace hmv
using namespace hmv;

extern "C" int hmv_peer_alloc(long long bytes, void** ptr_out, unsigned char* handle64_out) {
  HMV_REQUIRE(bytes > 0 && ptr_out && handle64_out, "hmv_peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e != cudaSuccess) return fail(HMV_E_CUDA, "hmv_peer_alloc: cudaMalloc(%lld): %s", bytes, cudaGetErrorString(e));
  e = cudaMemset(p, 0, (size_t)bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(HMV_E_CUDA, "hmv_peer_alloc: %s", cudaGetErrorString(e));
  }
  memcpy(handle64_out, &h, 64);
  *ptr_out = p;
  return HMV_OK;
}

extern "C" int hmv_peer_open(const unsigned char* handle64, void** ptr_out) {
  HMV_REQUIRE(handle64 && ptr_out, "hmv_peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(HMV_E_CUDA, "hmv_peer_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  }
  *ptr_out = p;
  return HMV_OK;
}

extern "C" int hmv_peer_close(void* ptr) {
  if (!ptr) return HMV_OK;
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(HMV_E_CUDA, "hmv_peer_close: %s", cudaGetErrorString(e)); }
  return HMV_OK;
}

extern "C" int hmv_peer_free(void* ptr) {
  if (!ptr) return HMV_OK;
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(HMV_E_CUDA, "hmv_peer_free: %s", cudaGetErrorString(e)); }
  return HMV_OK;
}

extern "C" int hmv_peer_scatter(int nrow, int ncol, int nsp, const double* const* a_h, const double* const* b_h,
                                int npeers, int rank, void* const* peer_buf_h, void* const* peer_flag_h,
                                long long row0, unsigned long long step, unsigned int* done_d, void* stream) {
  HMV_REQUIRE(nrow > 0 && nrow <= 65535 && ncol > 0 && nsp >= 1 && nsp <= 4 && a_h && peer_buf_h && peer_flag_h && done_d,
              "hmv_peer_scatter: bad arguments");
  HMV_REQUIRE(npeers >= 1 && npeers <= HMV_MAX_PEERS && rank >= 0 && rank < npeers && row0 >= 0 && step > 0,
              "hmv_peer_scatter: npeers=%d (max %d), rank=%d", npeers, HMV_MAX_PEERS, rank);
  PeerArgs q;
  memset(&q, 0, sizeof(q));
  for (int s = 0; s < nsp; ++s) {
    q.a[s] = a_h[s];
    q.b[s] = b_h ? b_h[s] : nullptr;
    HMV_REQUIRE(q.a[s], "hmv_peer_scatter: null table");
  }
  for (int p = 0; p < npeers; ++p) {
    q.buf[p] = (double*)peer_buf_h[p];
    q.flag[p] = (unsigned long long*)peer_flag_h[p];
    HMV_REQUIRE(q.buf[p] && q.flag[p], "hmv_peer_scatter: null peer pointer (rank %d)", p);
  }
  q.nrow = nrow; q.ncol = ncol; q.nsp = nsp; q.npeers = npeers; q.rank = rank; q.row0 = row0; q.step = step;
  q.done = done_d;
  dim3 grid(cdiv(ncol, 512), nrow);
  peer_scatter_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(q);
  return check_launch("peer_scatter_kernel");
}

extern "C" int hmv_peer_wait(const void* flags_d, int npeers, unsigned long long step, double timeout_s, int* status_d,
                             void* stream) {
  HMV_REQUIRE(flags_d && status_d && npeers >= 1 && npeers <= HMV_MAX_PEERS && timeout_s > 0.0, "hmv_peer_wait: bad arguments");
  static int khz = 0;                       // queried once: this attribute costs milliseconds per call
  if (khz <= 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev) != cudaSuccess || khz <= 0)
      khz = 1900000;
  }
  const long long cycles = (long long)(timeout_s * 1.0e3 * (double)khz);
  peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const unsigned long long*)flags_d, npeers, step, cycles, status_d);
  return check_launch("peer_wait_kernel");
}
