// S1 of SURVEY.md 2b: one-time COO -> CSR-by-destination build.  A stable LSD radix
// sort of the edges on their destination id (8-bit digits, only as many passes as
// log2(N) needs) so that, inside a row, edges keep their edge_index order - the
// aggregation kernels then sum in a fixed, run-to-run identical order.
#include "common.cuh"

namespace dgcn {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_CHUNK = RS_THREADS * RS_ITEMS;

__global__ void csr_init_kernel(const int64_t* __restrict__ edge_index, int64_t E, int N,
                                int32_t* __restrict__ keys, int32_t* __restrict__ vals, int32_t* __restrict__ deg) {
  int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int64_t d = edge_index[E + e];
  d = d < 0 ? 0 : (d >= N ? N - 1 : d);
  keys[e] = static_cast<int32_t>(d);
  vals[e] = static_cast<int32_t>(e);
  atomicAdd(&deg[d], 1);
}

__global__ void radix_hist_kernel(const int32_t* __restrict__ keys, int64_t E, int shift, int nblocks,
                                  int32_t* __restrict__ hist) {
  __shared__ int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = static_cast<int64_t>(blockIdx.x) * RS_CHUNK;
  for (int r = 0; r < RS_ITEMS; ++r) {
    int64_t i = base + r * RS_THREADS + threadIdx.x;
    if (i < E) atomicAdd(&h[(keys[i] >> shift) & 255], 1);
  }
  __syncthreads();
  hist[static_cast<int64_t>(threadIdx.x) * nblocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void radix_scatter_kernel(const int32_t* __restrict__ keys, const int32_t* __restrict__ vals, int64_t E,
                                     int shift, int nblocks, const int32_t* __restrict__ offs,
                                     int32_t* __restrict__ keys_out, int32_t* __restrict__ vals_out) {
  __shared__ int base[256];
  __shared__ int wcount[8][256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  base[tid] = offs[static_cast<int64_t>(tid) * nblocks + blockIdx.x];
  const int64_t cbase = static_cast<int64_t>(blockIdx.x) * RS_CHUNK;
  for (int r = 0; r < RS_ITEMS; ++r) {
    for (int w = 0; w < 8; ++w) wcount[w][tid] = 0;
    __syncthreads();
    const int64_t i = cbase + r * RS_THREADS + tid;
    const bool valid = i < E;
    int key = 0, val = 0, d = 0x7FFF0000 + lane;   // invalid lanes never match anyone
    if (valid) {
      key = keys[i];
      val = vals[i];
      d = (key >> shift) & 255;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank == 0) wcount[warp][d] = __popc(peers);
    __syncthreads();
    {
      int run = base[tid];
      for (int w = 0; w < 8; ++w) {
        int c = wcount[w][tid];
        wcount[w][tid] = run;
        run += c;
      }
      base[tid] = run;
    }
    __syncthreads();
    if (valid) {
      const int pos = wcount[warp][d] + rank;
      keys_out[pos] = key;
      vals_out[pos] = val;
    }
    __syncthreads();
  }
}

// exclusive scan, int32, arbitrary n: (1) per-1024 block scan + totals, (2) one block
// scans the totals sequentially in 1024-wide strips, (3) add back.
__global__ void scan_block_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, int64_t n,
                                  int32_t* __restrict__ totals) {
  __shared__ int s[1024];
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 1024 + threadIdx.x;
  int v = i < n ? in[i] : 0;
  s[threadIdx.x] = v;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
    __syncthreads();
    s[threadIdx.x] += t;
    __syncthreads();
  }
  if (i < n) out[i] = s[threadIdx.x] - v;
  if (threadIdx.x == 1023 && totals) totals[blockIdx.x] = s[1023];
}
__global__ void scan_totals_kernel(int32_t* __restrict__ totals, int64_t nb) {
  __shared__ int s[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t b0 = 0; b0 < nb; b0 += 1024) {
    const int64_t i = b0 + threadIdx.x;
    int v = i < nb ? totals[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nb) totals[i] = s[threadIdx.x] - v + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry += s[1023];
    __syncthreads();
  }
}
__global__ void scan_add_kernel(int32_t* __restrict__ out, int64_t n, const int32_t* __restrict__ totals) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 1024 + threadIdx.x;
  if (i < n) out[i] += totals[blockIdx.x];
}

static int exclusive_scan(const int32_t* in, int32_t* out, int64_t n, int32_t* totals, cudaStream_t stream) {
  const int64_t nb = ceil_div(n, 1024);
  scan_block_kernel<<<static_cast<unsigned>(nb), 1024, 0, stream>>>(in, out, n, totals);
  DGCN_LAUNCH_CHECK();
  if (nb > 1) {
    scan_totals_kernel<<<1, 1024, 0, stream>>>(totals, nb);
    DGCN_LAUNCH_CHECK();
    scan_add_kernel<<<static_cast<unsigned>(nb), 1024, 0, stream>>>(out, n, totals);
    DGCN_LAUNCH_CHECK();
  }
  return DGCN_OK;
}

__global__ void csr_fill_src_kernel(const int64_t* __restrict__ edge_index, const int32_t* __restrict__ eid,
                                    int64_t E, int N, int32_t* __restrict__ src) {
  int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int64_t s = edge_index[eid[e]];
  src[e] = static_cast<int32_t>(s < 0 ? 0 : s);   // sources may index a different row set than N (halo / explicit messages)
}

struct CsrPlan {
  int64_t nblocks, hist, scan_tot;
};
static CsrPlan csr_plan(int64_t N, int64_t E) {
  CsrPlan p;
  p.nblocks = ceil_div(E > 0 ? E : 1, RS_CHUNK);
  p.hist = 256 * p.nblocks;
  int64_t longest = p.hist > N + 1 ? p.hist : N + 1;
  p.scan_tot = ceil_div(longest, 1024) + 1;
  return p;
}

}  // namespace dgcn

using namespace dgcn;

extern "C" {

size_t dgcn_csr_build_workspace_bytes(int64_t N, int64_t E) {
  CsrPlan p = csr_plan(N, E);
  size_t b = 0;
  b += 3 * align_up(static_cast<size_t>(E > 0 ? E : 1) * 4, 256);   // keys x2, vals x1
  b += align_up(static_cast<size_t>(N + 1) * 4, 256);                // deg
  b += 2 * align_up(static_cast<size_t>(p.hist) * 4, 256);           // hist, offsets
  b += align_up(static_cast<size_t>(p.scan_tot) * 4, 256);
  return b + 256;
}

int dgcn_csr_build(const int64_t* edge_index, int64_t E, int64_t N, int32_t* rowptr, int32_t* src, int32_t* eid,
                   void* wsp, size_t ws_bytes, dgcn_stream_t stream_) {
  if (!rowptr || !src || !eid || N <= 0 || E < 0 || (E > 0 && !edge_index)) return DGCN_ERR_BAD_ARG;
  if (N >= (1ll << 31) - 1 || E >= (1ll << 31) - 1) return DGCN_ERR_UNSUPPORTED;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Workspace ws(wsp, ws_bytes);
  CsrPlan p = csr_plan(N, E);
  const size_t en = static_cast<size_t>(E > 0 ? E : 1);
  int32_t* keys_a = ws.take<int32_t>(en);
  int32_t* keys_b = ws.take<int32_t>(en);
  int32_t* vals_b = ws.take<int32_t>(en);
  int32_t* deg = ws.take<int32_t>(static_cast<size_t>(N + 1));
  int32_t* hist = ws.take<int32_t>(static_cast<size_t>(p.hist));
  int32_t* offs = ws.take<int32_t>(static_cast<size_t>(p.hist));
  int32_t* tot = ws.take<int32_t>(static_cast<size_t>(p.scan_tot));
  if (!ws.ok) return DGCN_ERR_WORKSPACE;
  DGCN_CUDA_TRY(cudaMemsetAsync(deg, 0, static_cast<size_t>(N + 1) * 4, stream));
  if (E > 0) {
    csr_init_kernel<<<static_cast<unsigned>(ceil_div(E, 256)), 256, 0, stream>>>(edge_index, E, static_cast<int>(N),
                                                                               keys_a, eid, deg);
    DGCN_LAUNCH_CHECK();
  }
  int rc = exclusive_scan(deg, rowptr, N + 1, tot, stream);
  if (rc != DGCN_OK) return rc;
  if (E == 0) return DGCN_OK;
  int bits = 1;
  while ((1ll << bits) < N) ++bits;
  int passes = (bits + 7) / 8;
  // ping-pong so that the final pass lands in (keys_?, eid)
  int32_t* kin = keys_a;
  int32_t* vin = eid;
  int32_t* kout = keys_b;
  int32_t* vout = vals_b;
  if (passes % 2 == 1) {   // odd number of passes: start from vals_b so the last write hits eid
    DGCN_CUDA_TRY(cudaMemcpyAsync(vals_b, eid, static_cast<size_t>(E) * 4, cudaMemcpyDeviceToDevice, stream));
    vin = vals_b;
    vout = eid;
  }
  for (int ps = 0; ps < passes; ++ps) {
    const int shift = ps * 8;
    radix_hist_kernel<<<static_cast<unsigned>(p.nblocks), RS_THREADS, 0, stream>>>(kin, E, shift,
                                                                                   static_cast<int>(p.nblocks), hist);
    DGCN_LAUNCH_CHECK();
    rc = exclusive_scan(hist, offs, p.hist, tot, stream);
    if (rc != DGCN_OK) return rc;
    radix_scatter_kernel<<<static_cast<unsigned>(p.nblocks), RS_THREADS, 0, stream>>>(
        kin, vin, E, shift, static_cast<int>(p.nblocks), offs, kout, vout);
    DGCN_LAUNCH_CHECK();
    int32_t* t = kin; kin = kout; kout = t;
    t = vin; vin = vout; vout = t;
  }
  // after the loop `vin` holds the sorted edge ids; by construction vin == eid
  csr_fill_src_kernel<<<static_cast<unsigned>(ceil_div(E, 256)), 256, 0, stream>>>(edge_index, eid, E,
                                                                                 static_cast<int>(N), src);
  DGCN_LAUNCH_CHECK();
  return DGCN_OK;
}

}  // extern "C"
