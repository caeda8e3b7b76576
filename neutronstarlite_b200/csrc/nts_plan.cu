// nts_gather_plan: a chunk's (offsets, indices, weight) arrays preprocessed ONCE for repeated aggregation, and the
// kernel that consumes them.  Same contraction as segment_gather_sum_kernel (nts_aggregate.cu) -
//
//   out[r,:] += sum_{e in [off[r], off[r+1])} in[row(e),:] * w[e]
//
// i.e. Cuda_Stream::Gather_By_Dst_From_Src / Gather_By_Src_From_Dst (cuda/ntsCUDAGraphOP.cu:157-281) - but with the
// three things the ncu captures of round 1 asked for (profiles/README.md: the kernel is bound by the L1 data stage
// every gathered byte crosses, and re-reads the feature matrix from DRAM 26 times):
//
//   1. source-slab bucketing.  Edges are regrouped by (slab of the gathered row, output row): slab s holds the
//      edges whose gathered row lies in rows [s*slab_rows, (s+1)*slab_rows) of the input matrix, sized so that one
//      slab of the matrix stays resident in the 126 MB L2.  One launch per slab, in stream order, so at any moment
//      the CTAs in flight gather from ONE slab; the output row of a (slab, row) segment is finished with a plain
//      read-modify-write exactly like the unbucketed kernel (launches never overlap, so no extra atomics).
//      The bucketing is a stable sort by (slab, row): inside a segment the edges keep the order of the reference
//      layout (core/PartitionedGraph.hpp:389-405), summation order per output element changes only in where the
//      partial sums of the slabs are added.
//   2. (row, weight) pairs.  The base / slot lookup is applied at plan time and row index and weight are stored
//      interleaved, so the per-edge broadcast read from shared memory is ONE 8-byte LDS instead of two 4-byte ones,
//      and the TMA bulk copy (cp.async.bulk -> SASS UBLKCP) stages one array per CTA instead of two.
//   3. 16-byte feature loads for every width.  Rows whose byte length is not a multiple of 16 (F = 602: 2408 B) are
//      copied once per call into a workspace with rows padded to a multiple of 4 floats (0.2 ms of HBM time at
//      config B), so the gather always uses float4 loads on 16-byte aligned rows and one warp covers up to 640
//      columns: 19 + 1 data-stage wavefronts per edge at F = 602 instead of 21.6 + 4.
//
// Plan construction (hand-written kernels + one CUB radix sort) replaces nothing in the reference: its chunks are
// built on the host by single-threaded loops (core/PartitionedGraph.hpp:324-420) and never re-bucketed.
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "nts_common.cuh"

struct nts_gather_plan {
  uint32_t n_rows = 0;         // output rows
  uint64_t n_edges = 0;
  uint32_t gather_rows = 0;    // rows of the gathered matrix
  int slabs = 1;
  uint32_t slab_rows = 0;
  uint2 *pairs = nullptr;      // [n_edges] {gathered row, weight bits}, slab-major, then output row, then original order
  uint32_t *voff = nullptr;    // [slabs * n_rows + 1] offsets of the (slab, row) segments
  std::vector<uint64_t> slab_edge; // [slabs + 1] host copy of voff[s * n_rows]
  float *workspace = nullptr;  // padded copy of the input when its rows are not 16-byte multiples / aligned
  size_t workspace_floats = 0;
  int last_grid = 0, last_launches = 0, last_k = 0, last_u = 0, last_outv = 0;
  float tuned_ms = 0.f;        // nts_gather_plan_create_tuned: time of the winning candidate
};

namespace nts {

static int g_plan_u = 0, g_plan_minb = 0, g_plan_q = 0, g_plan_variant = 0; // measurement hooks (NTS_PLAN_TUNE="U,MINB[,Q]"), 0 = default

// ---- plan construction kernels -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t plan_find_row(const uint32_t *__restrict__ off, uint32_t n_rows, uint32_t e) {
  uint32_t lo = 0, hi = n_rows;
  while (hi - lo > 1) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(off + mid) <= e)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

// key[e] = slab(e) * n_rows + row(e), val[e] = e
__global__ void plan_keys_kernel(const uint32_t *__restrict__ off, const uint32_t *__restrict__ idx,
                                 const uint32_t *__restrict__ slot_of, uint32_t base, uint32_t n_rows, uint32_t n_edges,
                                 uint32_t slab_rows, uint32_t slabs, uint32_t *__restrict__ key,
                                 uint32_t *__restrict__ val) {
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t r = plan_find_row(off, n_rows, (uint32_t)e);
    const uint32_t id = __ldg(idx + e);
    const uint32_t g = slot_of ? __ldg(slot_of + id) : id - base;
    uint32_t s = g / slab_rows;
    if (s >= slabs)
      s = slabs - 1;
    key[e] = s * n_rows + r;
    val[e] = (uint32_t)e;
  }
}

// pairs[i] = {row(perm[i]), w[perm[i]]}   (perm == nullptr: identity)
__global__ void plan_pairs_kernel(const uint32_t *__restrict__ perm, const uint32_t *__restrict__ idx,
                                  const float *__restrict__ w, const uint32_t *__restrict__ slot_of, uint32_t base,
                                  uint32_t n_edges, uint2 *__restrict__ pairs) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_edges; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t e = perm ? __ldg(perm + i) : (uint32_t)i;
    const uint32_t id = __ldg(idx + e);
    const uint32_t g = slot_of ? __ldg(slot_of + id) : id - base;
    const float wt = w ? __ldg(w + e) : 1.f;
    pairs[i] = make_uint2(g, __float_as_uint(wt));
  }
}

// Multi-part variants (several chunks merged into one plan): part-local row r -> output row row_add + r, mapped index
// g -> index_add + g; edge e of the part is global edge e_off + e.
__global__ void plan_part_keys_kernel(const uint32_t *__restrict__ off, const uint32_t *__restrict__ idx,
                                      const uint32_t *__restrict__ slot_of, const float *__restrict__ w, uint32_t base,
                                      uint32_t index_add, uint32_t n_rows_part, uint32_t row_add, uint32_t n_edges,
                                      uint32_t e_off, uint32_t n_rows_total, uint32_t slab_rows, uint32_t slabs,
                                      uint32_t *__restrict__ key, uint32_t *__restrict__ val,
                                      uint2 *__restrict__ unsorted) {
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t r = plan_find_row(off, n_rows_part, (uint32_t)e) + row_add;
    const uint32_t id = __ldg(idx + e);
    const uint32_t g = (slot_of ? __ldg(slot_of + id) : id - base) + index_add;
    uint32_t s = g / slab_rows;
    if (s >= slabs)
      s = slabs - 1;
    key[e_off + e] = s * n_rows_total + r;
    val[e_off + e] = e_off + (uint32_t)e;
    unsorted[e_off + e] = make_uint2(g, __float_as_uint(w ? __ldg(w + e) : 1.f));
  }
}
__global__ void plan_permute_pairs_kernel(const uint32_t *__restrict__ perm, const uint2 *__restrict__ unsorted,
                                          uint32_t n_edges, uint2 *__restrict__ pairs) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_edges; i += (uint64_t)gridDim.x * blockDim.x)
    pairs[i] = unsorted[__ldg(perm + i)];
}

// voff[k] = number of sorted keys < k, k in [0, n_keys]
__global__ void plan_offsets_kernel(const uint32_t *__restrict__ sorted_key, uint32_t n_edges, uint32_t n_keys,
                                    uint32_t *__restrict__ voff) {
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k <= n_keys; k += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t lo = 0, hi = n_edges; // first position with key >= k
    while (lo < hi) {
      uint32_t mid = lo + ((hi - lo) >> 1);
      if (__ldg(sorted_key + mid) < (uint32_t)k)
        lo = mid + 1;
      else
        hi = mid;
    }
    voff[k] = lo;
  }
}

// dst[r, 0:ld] = {src[r, 0:F], 0...}   (ld = F rounded up to a multiple of 4; one warp per row piece)
__global__ void pad_rows_kernel(const float *__restrict__ src, float *__restrict__ dst, uint32_t n_rows, uint32_t F,
                                uint32_t ld) {
  const uint64_t total = (uint64_t)n_rows * ld;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = i / ld;
    const uint32_t c = (uint32_t)(i - r * ld);
    dst[i] = c < F ? __ldg(src + r * F + c) : 0.f;
  }
}

// ---- the aggregation kernel ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t p_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void p_mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(p_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void p_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(p_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void p_mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile("{\n\t"
               ".reg .pred p;\n\t"
               "WAIT_%=:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra DONE_%=;\n\t"
               "bra WAIT_%=;\n\t"
               "DONE_%=:\n\t"
               "}" ::"r"(p_smem_u32(bar)),
               "r"(parity)
               : "memory");
}
__device__ __forceinline__ void p_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   p_smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(p_smem_u32(bar))
               : "memory");
}

constexpr int kPlanWarps = 8;

// Output columns 4c .. 4c+3 of one accumulator chunk, OUTV floats per store (the output keeps the caller's row
// stride F, so its rows are 16-byte aligned only when F % 4 == 0).
template <int OUTV, bool ATOMIC>
__device__ __forceinline__ void flush_chunk(float *__restrict__ orow, uint32_t col, uint32_t F, float4 a) {
  if constexpr (OUTV == 4) {
    float4 *p = reinterpret_cast<float4 *>(orow + col);
    if constexpr (ATOMIC) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w)
                   : "memory");
    } else {
      float4 o = *p;
      o.x += a.x, o.y += a.y, o.z += a.z, o.w += a.w;
      *p = o;
    }
  } else if constexpr (OUTV == 2) {
    float2 *p = reinterpret_cast<float2 *>(orow + col);
    if constexpr (ATOMIC) {
      asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a.x), "f"(a.y) : "memory");
      if (col + 2 < F)
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p + 1), "f"(a.z), "f"(a.w) : "memory");
    } else {
      float2 o = p[0];
      o.x += a.x, o.y += a.y;
      p[0] = o;
      if (col + 2 < F) {
        float2 o1 = p[1];
        o1.x += a.z, o1.y += a.w;
        p[1] = o1;
      }
    }
  } else {
    const float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (col + i < F) {
        if constexpr (ATOMIC)
          atomicAdd(orow + col + i, v[i]);
        else
          orow[col + i] += v[i];
      }
  }
}

// K    : float4 chunks per lane per column tile (a tile covers K*128 floats)
// U    : edges whose K loads are issued before any FMA (U*K independent 16-byte loads per lane)
// OUTV : floats per output store (4 when F % 4 == 0 and the output is 16-byte aligned, else 2 or 1)
// MINB : __launch_bounds__ minimum CTAs per SM
// G    : virtual warps per warp.  Rows of at most 16 / 8 float4 (F <= 64 / 32) would leave half / three quarters of
//        the lanes idle, so a warp is split into G independent groups of 32/G lanes, each with its own edge quantum,
//        row bookkeeping and accumulators (the kernel has no warp-wide shuffles: all state is per lane already).
// Warp g owns the edge quantum [e_begin + q*Q, ...) of column tile t (g = q*tiles + t); the (row, weight) pairs of the
// CTA's edge span are staged in shared memory by one cp.async.bulk, completion on an mbarrier.
template <int K, int U, int OUTV, int MINB, int G = 1>
__global__ void __launch_bounds__(kPlanWarps * 32, MINB)
    planned_gather_sum_kernel(const float4 *__restrict__ in, uint32_t ld4, float *__restrict__ out, uint32_t F,
                              const uint2 *__restrict__ pairs, const uint32_t *__restrict__ off, uint32_t n_rows,
                              uint32_t e_begin, uint32_t e_end, uint32_t Q, uint32_t tiles, uint32_t tile_vecs) {
  static_assert(G == 1 || K == 1, "virtual warps are for rows narrower than a warp");
  constexpr uint32_t GS = 32 / G;                       // lanes per virtual warp
  const uint32_t lane = threadIdx.x & (GS - 1);         // lane within the virtual warp
  const uint32_t vwarp_in_block = threadIdx.x / GS;
  const uint64_t gwarp = (uint64_t)blockIdx.x * (kPlanWarps * G) + vwarp_in_block;
  const uint32_t tile = (uint32_t)(gwarp % tiles);
  const uint64_t q = gwarp / tiles;
  const uint64_t e0_64 = e_begin + q * (uint64_t)Q;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
  uint2 *s_pair = reinterpret_cast<uint2 *>(smem_raw + 16);
  uint32_t cta_e_base = 0, bulk_bytes = 0;
  {
    const uint64_t cta_w0 = (uint64_t)blockIdx.x * (kPlanWarps * G);
    const uint64_t first_q = cta_w0 / tiles;
    const uint64_t last_q = (cta_w0 + kPlanWarps * G - 1) / tiles;
    uint64_t ce0 = e_begin + first_q * (uint64_t)Q;
    uint64_t ce1 = e_begin + (last_q + 1) * (uint64_t)Q;
    if (ce1 > e_end)
      ce1 = e_end;
    if (ce0 < ce1) {
      cta_e_base = (uint32_t)(ce0 & ~1ull); // 16-byte aligned start (8-byte elements, array 16-byte aligned)
      const uint32_t n_el = (uint32_t)(ce1 - cta_e_base);
      bulk_bytes = (n_el * 8u) & ~15u;      // whole 16-byte units through the bulk engine, a last odd element by a plain load
      if (threadIdx.x == 0) {
        p_mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (n_el & 1u)
          s_pair[n_el - 1] = __ldg(pairs + cta_e_base + n_el - 1);
      }
      __syncthreads();
      if (threadIdx.x == 0 && bulk_bytes) {
        p_mbar_expect_tx(bar, bulk_bytes);
        p_bulk_g2s(s_pair, pairs + cta_e_base, bulk_bytes, bar);
      }
    }
  }
  if (e0_64 >= e_end)
    return;
  const uint32_t e0 = (uint32_t)e0_64;
  const uint32_t e1 = (e0_64 + Q < e_end) ? e0 + Q : e_end;

  const uint32_t c0 = tile * tile_vecs + lane; // first float4 column of this lane
  bool act[K];
#pragma unroll
  for (int k = 0; k < K; k++)
    act[k] = (k * GS + lane) < tile_vecs && (c0 + k * GS) < ld4;

  uint32_t row = plan_find_row(off, n_rows, e0);
  uint32_t row_end = __ldg(off + row + 1);
  bool row_started_inside = __ldg(off + row) >= e0;

  float4 acc[K];
#pragma unroll
  for (int k = 0; k < K; k++)
    acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);

  auto flush = [&](bool whole) {
    float *orow = out + (size_t)row * F;
#pragma unroll
    for (int k = 0; k < K; k++) {
      const uint32_t col = (c0 + k * GS) * 4;
      if (act[k] && col < F) {
        if (whole)
          flush_chunk<OUTV, false>(orow, col, F, acc[k]);
        else
          flush_chunk<OUTV, true>(orow, col, F, acc[k]);
      }
      acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto advance = [&](uint32_t ee) {
    flush(row_started_inside);
    do {
      row++;
      row_end = __ldg(off + row + 1);
    } while (ee >= row_end);
    row_started_inside = true;
  };

  if (bulk_bytes)
    p_mbar_wait(bar, 0);

  const uint2 *sp = s_pair - cta_e_base;
  uint32_t e = e0;
  for (; e + U <= e1; e += U) {
    float4 v[U][K];
    float wu[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint2 pr = sp[e + u];
      wu[u] = __uint_as_float(pr.y);
      const float4 *p = in + (size_t)pr.x * ld4 + c0;
#pragma unroll
      for (int k = 0; k < K; k++)
        if (act[k])
          v[u][k] = __ldg(p + k * GS);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (e + u >= row_end)
        advance(e + u);
#pragma unroll
      for (int k = 0; k < K; k++)
        if (act[k]) {
          acc[k].x = fmaf(wu[u], v[u][k].x, acc[k].x);
          acc[k].y = fmaf(wu[u], v[u][k].y, acc[k].y);
          acc[k].z = fmaf(wu[u], v[u][k].z, acc[k].z);
          acc[k].w = fmaf(wu[u], v[u][k].w, acc[k].w);
        }
    }
  }
  for (; e < e1; e++) {
    const uint2 pr = sp[e];
    const float wj = __uint_as_float(pr.y);
    const float4 *p = in + (size_t)pr.x * ld4 + c0;
    float4 v1[K];
#pragma unroll
    for (int k = 0; k < K; k++)
      if (act[k])
        v1[k] = __ldg(p + k * GS);
    if (e >= row_end)
      advance(e);
#pragma unroll
    for (int k = 0; k < K; k++)
      if (act[k]) {
        acc[k].x = fmaf(wj, v1[k].x, acc[k].x);
        acc[k].y = fmaf(wj, v1[k].y, acc[k].y);
        acc[k].z = fmaf(wj, v1[k].z, acc[k].z);
        acc[k].w = fmaf(wj, v1[k].w, acc[k].w);
      }
  }
  flush(row_started_inside && row_end <= e1);
}

struct PlanShape;
// ---- experiment the north star asks for: feature ROWS staged in shared memory by TMA ----------------------------------
// Same work split and row bookkeeping, but the gathered row never passes through registers on its way in: lane 0 of
// each warp issues one cp.async.bulk (SASS UBLKCP) per edge that copies the row's tile (16-byte aligned thanks to the
// padded workspace) into a per-warp ring of STAGES shared-memory buffers, completion on one mbarrier per stage; the
// warp waits, reads its chunks with 16-byte LDS, accumulates, and re-arms the stage for edge e + STAGES.
// Kept as variant 1 of nts_gather_plan_set_variant for measurement (profiles/): bulk copies bypass L1, which serves
// the hub rows of a skewed graph, and every gathered byte still crosses the shared-memory data stage once.
template <int K, int STAGES, int OUTV, int MINB>
__global__ void __launch_bounds__(kPlanWarps * 32, MINB)
    planned_gather_sum_tma_kernel(const float4 *__restrict__ in, uint32_t ld4, float *__restrict__ out, uint32_t F,
                                  const uint2 *__restrict__ pairs, const uint32_t *__restrict__ off, uint32_t n_rows,
                                  uint32_t e_begin, uint32_t e_end, uint32_t Q, uint32_t tiles, uint32_t tile_vecs,
                                  uint32_t pair_bytes) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp_in_block = threadIdx.x >> 5;
  const uint64_t gwarp = (uint64_t)blockIdx.x * kPlanWarps + warp_in_block;
  const uint32_t tile = (uint32_t)(gwarp % tiles);
  const uint64_t q = gwarp / tiles;
  const uint64_t e0_64 = e_begin + q * (uint64_t)Q;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
  uint2 *s_pair = reinterpret_cast<uint2 *>(smem_raw + 16);
  // per-warp ring: STAGES barriers, then STAGES row buffers of tile_vecs float4
  const uint32_t ring_bytes = 8u * STAGES + 8u * (STAGES & 1) + STAGES * tile_vecs * 16u;
  unsigned char *ring = smem_raw + 16 + pair_bytes + (size_t)warp_in_block * ring_bytes;
  uint64_t *sbar = reinterpret_cast<uint64_t *>(ring);
  float4 *sbuf = reinterpret_cast<float4 *>(ring + 8u * STAGES + 8u * (STAGES & 1));
  uint32_t cta_e_base = 0, bulk_bytes = 0;
  {
    const uint64_t cta_w0 = (uint64_t)blockIdx.x * kPlanWarps;
    const uint64_t first_q = cta_w0 / tiles;
    const uint64_t last_q = (cta_w0 + kPlanWarps - 1) / tiles;
    uint64_t ce0 = e_begin + first_q * (uint64_t)Q;
    uint64_t ce1 = e_begin + (last_q + 1) * (uint64_t)Q;
    if (ce1 > e_end)
      ce1 = e_end;
    if (lane == 0)
      for (int s = 0; s < STAGES; s++)
        p_mbar_init(sbar + s, 1);
    if (ce0 < ce1) {
      cta_e_base = (uint32_t)(ce0 & ~1ull);
      const uint32_t n_el = (uint32_t)(ce1 - cta_e_base);
      bulk_bytes = (n_el * 8u) & ~15u;
      if (threadIdx.x == 0) {
        p_mbar_init(bar, 1);
        if (n_el & 1u)
          s_pair[n_el - 1] = __ldg(pairs + cta_e_base + n_el - 1);
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0 && bulk_bytes) {
      p_mbar_expect_tx(bar, bulk_bytes);
      p_bulk_g2s(s_pair, pairs + cta_e_base, bulk_bytes, bar);
    }
  }
  if (e0_64 >= e_end)
    return;
  const uint32_t e0 = (uint32_t)e0_64;
  const uint32_t e1 = (e0_64 + Q < e_end) ? e0 + Q : e_end;
  const uint32_t c0 = tile * tile_vecs + lane;
  const uint32_t my_vecs = min(tile_vecs, ld4 - tile * tile_vecs); // float4 of this tile that exist in the row
  const uint32_t row_bytes = my_vecs * 16u;
  bool act[K];
#pragma unroll
  for (int k = 0; k < K; k++)
    act[k] = (k * 32 + lane) < my_vecs;

  uint32_t row = plan_find_row(off, n_rows, e0);
  uint32_t row_end = __ldg(off + row + 1);
  bool row_started_inside = __ldg(off + row) >= e0;
  float4 acc[K];
#pragma unroll
  for (int k = 0; k < K; k++)
    acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  auto flush = [&](bool whole) {
    float *orow = out + (size_t)row * F;
#pragma unroll
    for (int k = 0; k < K; k++) {
      const uint32_t col = (c0 + k * 32) * 4;
      if (act[k] && col < F) {
        if (whole)
          flush_chunk<OUTV, false>(orow, col, F, acc[k]);
        else
          flush_chunk<OUTV, true>(orow, col, F, acc[k]);
      }
      acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto advance = [&](uint32_t ee) {
    flush(row_started_inside);
    do {
      row++;
      row_end = __ldg(off + row + 1);
    } while (ee >= row_end);
    row_started_inside = true;
  };
  if (bulk_bytes)
    p_mbar_wait(bar, 0);
  const uint2 *sp = s_pair - cta_e_base;
  auto issue = [&](uint32_t e, uint32_t stage) { // lane 0 only
    const uint32_t src_row = sp[e].x;
    p_mbar_expect_tx(sbar + stage, row_bytes);
    p_bulk_g2s(sbuf + (size_t)stage * tile_vecs, in + (size_t)src_row * ld4 + tile * tile_vecs, row_bytes, sbar + stage);
  };
  if (lane == 0)
    for (uint32_t s = 0; s < STAGES && e0 + s < e1; s++)
      issue(e0 + s, s);
  uint32_t stage = 0, parity = 0;
  for (uint32_t e = e0; e < e1; e++) {
    const float w = __uint_as_float(sp[e].y);
    p_mbar_wait(sbar + stage, parity);
    const float4 *b = sbuf + (size_t)stage * tile_vecs + lane;
    float4 v[K];
#pragma unroll
    for (int k = 0; k < K; k++)
      if (act[k])
        v[k] = b[k * 32];
    if (e >= row_end)
      advance(e);
#pragma unroll
    for (int k = 0; k < K; k++)
      if (act[k]) {
        acc[k].x = fmaf(w, v[k].x, acc[k].x);
        acc[k].y = fmaf(w, v[k].y, acc[k].y);
        acc[k].z = fmaf(w, v[k].z, acc[k].z);
        acc[k].w = fmaf(w, v[k].w, acc[k].w);
      }
    __syncwarp(); // every lane has consumed the stage (its values are in registers and used): it may be overwritten
    if (lane == 0 && e + STAGES < e1)
      issue(e + STAGES, stage);
    if (++stage == STAGES) {
      stage = 0;
      parity ^= 1u;
    }
  }
  flush(row_started_inside && row_end <= e1);
}

template <int K, int STAGES, int OUTV, int MINB>
static int launch_planned_tma(nts_gather_plan *pl, const PlanShape &sh, const float4 *in, uint32_t ld4, float *out,
                              uint32_t F, uint32_t Q, cudaStream_t st);

struct PlanShape {
  int k, u, outv, minb, g;
  uint32_t tiles, tile_vecs;
};

template <int K, int U, int OUTV, int MINB, int G = 1>
static int launch_planned(nts_gather_plan *pl, const PlanShape &sh, const float4 *in, uint32_t ld4, float *out,
                          uint32_t F, uint32_t Q, cudaStream_t st) {
  auto kern = planned_gather_sum_kernel<K, U, OUTV, MINB, G>;
  const size_t smem = 16 + ((size_t)kPlanWarps * G * Q + 4) * 8;
  NTS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pl->last_launches = 0;
  for (int s = 0; s < pl->slabs; s++) {
    const uint64_t eb = pl->slab_edge[s], ee = pl->slab_edge[s + 1];
    if (ee <= eb)
      continue;
    const uint64_t quanta = (ee - eb + Q - 1) / Q;
    const uint64_t blocks = (quanta * sh.tiles + kPlanWarps * G - 1) / (kPlanWarps * G);
    NTS_ARG_CHECK(blocks <= 0x7fffffffull, "aggregation grid too large");
    kern<<<(unsigned)blocks, kPlanWarps * 32, smem, st>>>(in, ld4, out, F, pl->pairs, pl->voff + (size_t)s * pl->n_rows,
                                                          pl->n_rows, (uint32_t)eb, (uint32_t)ee, Q, sh.tiles,
                                                          sh.tile_vecs);
    NTS_LAUNCH_CHECK();
    pl->last_grid = (int)blocks;
    pl->last_launches++;
  }
  return 0;
}

template <int K, int STAGES, int OUTV, int MINB>
static int launch_planned_tma(nts_gather_plan *pl, const PlanShape &sh, const float4 *in, uint32_t ld4, float *out,
                              uint32_t F, uint32_t Q, cudaStream_t st) {
  auto kern = planned_gather_sum_tma_kernel<K, STAGES, OUTV, MINB>;
  const uint32_t pair_bytes = (kPlanWarps * Q + 4) * 8;
  const uint32_t ring_bytes = 8u * STAGES + 8u * (STAGES & 1) + STAGES * sh.tile_vecs * 16u;
  const size_t smem = 16 + pair_bytes + (size_t)kPlanWarps * ring_bytes;
  NTS_ARG_CHECK(smem <= 227 * 1024, "row-staging ring does not fit shared memory");
  NTS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pl->last_launches = 0;
  for (int s = 0; s < pl->slabs; s++) {
    const uint64_t eb = pl->slab_edge[s], ee = pl->slab_edge[s + 1];
    if (ee <= eb)
      continue;
    const uint64_t quanta = (ee - eb + Q - 1) / Q;
    const uint64_t blocks = (quanta * sh.tiles + kPlanWarps - 1) / kPlanWarps;
    NTS_ARG_CHECK(blocks <= 0x7fffffffull, "aggregation grid too large");
    kern<<<(unsigned)blocks, kPlanWarps * 32, smem, st>>>(in, ld4, out, F, pl->pairs, pl->voff + (size_t)s * pl->n_rows,
                                                          pl->n_rows, (uint32_t)eb, (uint32_t)ee, Q, sh.tiles,
                                                          sh.tile_vecs, pair_bytes);
    NTS_LAUNCH_CHECK();
    pl->last_grid = (int)blocks;
    pl->last_launches++;
  }
  return 0;
}

#define NTS_PLAN_TMA_CASE(K_, S_, B_)                                                                               \
  if (sh.k == K_ && stages == S_ && sh.minb == B_) {                                                                \
    if (sh.outv == 4)                                                                                               \
      return launch_planned_tma<K_, S_, 4, B_>(pl, sh, in4, ld4, out, F, Q, st);                                    \
    if (sh.outv == 2)                                                                                               \
      return launch_planned_tma<K_, S_, 2, B_>(pl, sh, in4, ld4, out, F, Q, st);                                    \
    return launch_planned_tma<K_, S_, 1, B_>(pl, sh, in4, ld4, out, F, Q, st);                                      \
  }

#define NTS_PLAN_CASE_G(U_, B_, G_)                                                                                 \
  if (sh.k == 1 && sh.u == U_ && sh.minb == B_ && sh.g == G_) {                                                     \
    if (sh.outv == 4)                                                                                               \
      return launch_planned<1, U_, 4, B_, G_>(pl, sh, in4, ld4, out, F, Q, st);                                     \
    if (sh.outv == 2)                                                                                               \
      return launch_planned<1, U_, 2, B_, G_>(pl, sh, in4, ld4, out, F, Q, st);                                     \
    return launch_planned<1, U_, 1, B_, G_>(pl, sh, in4, ld4, out, F, Q, st);                                       \
  }

#define NTS_PLAN_CASE(K_, U_, B_)                                                                                   \
  if (sh.k == K_ && sh.u == U_ && sh.minb == B_) {                                                                  \
    if (sh.outv == 4)                                                                                               \
      return launch_planned<K_, U_, 4, B_>(pl, sh, in4, ld4, out, F, Q, st);                                         \
    if (sh.outv == 2)                                                                                               \
      return launch_planned<K_, U_, 2, B_>(pl, sh, in4, ld4, out, F, Q, st);                                         \
    return launch_planned<K_, U_, 1, B_>(pl, sh, in4, ld4, out, F, Q, st);                                           \
  }

static int run_plan(nts_gather_plan *pl, const float *input, float *output, uint32_t F, cudaStream_t st) {
  if (pl->n_rows == 0 || pl->n_edges == 0 || F == 0)
    return 0;
  NTS_ARG_CHECK(input && output, "null feature pointer");
  // 16-byte loads need 16-byte aligned rows: otherwise gather from a zero-padded copy (ld = F rounded up to 4)
  const uint32_t ld = (F + 3u) & ~3u;
  const float *in = input;
  if (ld != F || !aligned_to(input, 16)) {
    const size_t need = (size_t)pl->gather_rows * ld;
    if (need > pl->workspace_floats) {
      if (pl->workspace)
        NTS_CUDA_OK(cudaFree(pl->workspace));
      pl->workspace = nullptr;
      NTS_CUDA_OK(cudaMalloc(reinterpret_cast<void **>(&pl->workspace), need * sizeof(float)));
      pl->workspace_floats = need;
    }
    const uint64_t total = (uint64_t)pl->gather_rows * ld;
    const unsigned blocks = (unsigned)std::min<uint64_t>((total + 255) / 256, (uint64_t)sm_count() * 32);
    pad_rows_kernel<<<blocks, 256, 0, st>>>(input, pl->workspace, pl->gather_rows, F, ld);
    NTS_LAUNCH_CHECK();
    in = pl->workspace;
  }
  PlanShape sh;
  const uint32_t ld4 = ld / 4;
  const uint32_t chunks = (ld4 + 31) / 32;
  const uint32_t kmax = 5;
  sh.tiles = (chunks + kmax - 1) / kmax;
  sh.tile_vecs = (ld4 + sh.tiles - 1) / sh.tiles;
  sh.k = (int)((sh.tile_vecs + 31) / 32);
  sh.tiles = (ld4 + sh.tile_vecs - 1) / sh.tile_vecs;
  sh.outv = (F % 4 == 0 && aligned_to(output, 16)) ? 4 : ((F % 2 == 0 && aligned_to(output, 8)) ? 2 : 1);
  sh.g = ld4 <= 8 ? 4 : (ld4 <= 16 ? 2 : 1); // rows narrower than half / a quarter of a warp: virtual warps
  if (g_plan_variant == 1 || getenv("NTS_PLAN_NO_SUBWARP"))
    sh.g = 1;
  // (U, min CTAs/SM): U*K 16-byte loads in flight per lane
  // defaults = the largest U that compiles without spills at the occupancy point (ptxas -v); measured points for
  // the headline shapes in profiles/ (tools/k1_sweep.py)
  // (k = 5, F = 602: U=4 at 2 CTAs/SM 12.75 ms vs U=2 13.9 / U=1 at 3 CTAs 13.5; k = 1, F = 128: U=4 at 4 CTAs 2.90 ms
  // vs U=8 at 3 CTAs 3.07 - profiles/k1_sweep_r2a.jsonl)
  sh.minb = sh.k >= 3 ? 2 : (sh.k == 2 ? 3 : 4);
  sh.u = sh.k == 4 ? 2 : 4;
  {
    static int env_read = 0;
    if (!env_read) {
      env_read = 1;
      if (const char *t = getenv("NTS_PLAN_TUNE"))
        sscanf(t, "%d,%d,%d", &g_plan_u, &g_plan_minb, &g_plan_q);
    }
    if (g_plan_u > 0)
      sh.u = g_plan_u;
    if (g_plan_minb > 0)
      sh.minb = g_plan_minb;
  }
  uint32_t Q = g_plan_q > 0 ? (uint32_t)g_plan_q : 512u / sh.g; // the CTA's staged span stays 8 * 512 pairs
  if (g_plan_q <= 0) { // shrink for small inputs so every slab launch still fills the SMs
    const uint64_t per_slab = pl->n_edges / (uint64_t)pl->slabs + 1;
    const uint64_t want_warps = (uint64_t)sm_count() * 64 * sh.g;
    while (Q > 32 && ((per_slab + Q - 1) / Q) * sh.tiles < want_warps)
      Q >>= 1;
  }
  Q = (Q + 31u) & ~31u;
  if (Q * sh.g > 1024)
    Q = (1024 / sh.g) & ~31u;
  pl->last_k = sh.k, pl->last_u = sh.u, pl->last_outv = sh.outv;
  const float4 *in4 = reinterpret_cast<const float4 *>(in);
  float *out = output;
  if (g_plan_variant == 1) { // TMA row staging (measurement variant): U = ring depth
    const int stages = sh.u >= 8 ? 8 : (sh.u >= 4 ? 4 : 2);
    sh.minb = g_plan_minb > 0 ? g_plan_minb : (sh.k >= 4 ? 2 : 4);
    NTS_PLAN_TMA_CASE(1, 8, 4)
    NTS_PLAN_TMA_CASE(1, 4, 4)
    NTS_PLAN_TMA_CASE(1, 8, 3)
    NTS_PLAN_TMA_CASE(2, 4, 3)
    NTS_PLAN_TMA_CASE(2, 4, 2)
    NTS_PLAN_TMA_CASE(3, 4, 2)
    NTS_PLAN_TMA_CASE(4, 4, 2)
    NTS_PLAN_TMA_CASE(4, 2, 2)
    NTS_PLAN_TMA_CASE(5, 4, 2)
    NTS_PLAN_TMA_CASE(5, 2, 2)
    NTS_PLAN_TMA_CASE(5, 4, 1)
    NTS_PLAN_TMA_CASE(5, 8, 1)
    return fail(-1, "no TMA row-staging instantiation for this (chunks, stages, occupancy) point", __FILE__, __LINE__);
  }
  NTS_PLAN_CASE_G(4, 4, 2)
  NTS_PLAN_CASE_G(4, 4, 4)
  NTS_PLAN_CASE_G(8, 3, 2)
  NTS_PLAN_CASE_G(8, 3, 4)
  if (sh.g != 1)
    return fail(-1, "no virtual-warp instantiation for this (U, occupancy) point", __FILE__, __LINE__);
  NTS_PLAN_CASE(1, 8, 4)
  NTS_PLAN_CASE(1, 4, 4)
  NTS_PLAN_CASE(1, 8, 3)
  NTS_PLAN_CASE(1, 16, 2)
  NTS_PLAN_CASE(2, 4, 3)
  NTS_PLAN_CASE(2, 8, 2)
  NTS_PLAN_CASE(3, 4, 3)
  NTS_PLAN_CASE(3, 4, 2)
  NTS_PLAN_CASE(4, 2, 2)
  NTS_PLAN_CASE(4, 4, 2)
  NTS_PLAN_CASE(5, 2, 2)
  NTS_PLAN_CASE(5, 1, 3)
  NTS_PLAN_CASE(5, 2, 3)
  NTS_PLAN_CASE(5, 4, 1)
  NTS_PLAN_CASE(5, 4, 2)
  return fail(-1, "no planned-aggregation instantiation for this (chunks, U, occupancy) point", __FILE__, __LINE__);
}

} // namespace nts

using namespace nts;

extern "C" {

int nts_gather_plan_pick_slabs(nts_vid_t gather_rows, uint64_t n_edges, nts_vid_t n_rows, nts_vid_t feature_size,
                               uint64_t l2_budget_bytes) {
  if (!l2_budget_bytes)
    l2_budget_bytes = 40ull << 20; // a slab that stays resident next to the streamed outputs and index tiles
  const uint64_t bytes = (uint64_t)gather_rows * ((feature_size + 3u) & ~3u) * 4ull;
  uint64_t s = (bytes + l2_budget_bytes - 1) / l2_budget_bytes;
  // every (slab, row) segment costs one read-modify-write of the output row: keep >= 16 edges per segment on average
  const uint64_t by_degree = n_rows ? n_edges / ((uint64_t)n_rows * 16ull) : 1;
  if (s > by_degree)
    s = by_degree;
  if (s > 64)
    s = 64;
  return s < 1 ? 1 : (int)s;
}

nts_gather_plan *nts_gather_plan_create(const nts_vid_t *offsets, const nts_vid_t *indices, const float *weight,
                                        const nts_vid_t *slot_of, nts_vid_t index_base, nts_vid_t n_rows,
                                        uint64_t n_edges, nts_vid_t gather_rows, int n_slabs, void *stream) {
  auto bad = [](const char *m) -> nts_gather_plan * {
    fail(-1, m, __FILE__, __LINE__);
    return nullptr;
  };
  if (n_edges >= 0xffffffffull)
    return bad("chunk edge count must fit uint32 offsets");
  if (n_rows && n_edges && !(offsets && indices))
    return bad("null graph array");
  if (n_slabs < 1)
    n_slabs = 1;
  if (gather_rows == 0)
    n_slabs = 1;
  if ((uint64_t)n_slabs * n_rows >= 0xffffffffull)
    return bad("slabs * rows must fit 32-bit segment keys");
  cudaStream_t st = as_stream(stream);
  nts_gather_plan *pl = new nts_gather_plan();
  pl->n_rows = n_rows;
  pl->n_edges = n_edges;
  pl->gather_rows = gather_rows;
  pl->slabs = n_slabs;
  pl->slab_rows = gather_rows ? (gather_rows + n_slabs - 1) / n_slabs : 1;
  pl->slab_edge.assign(n_slabs + 1, 0);
  if (n_rows == 0 || n_edges == 0)
    return pl;
  const uint32_t E = (uint32_t)n_edges;
  const uint32_t n_keys = (uint32_t)n_slabs * n_rows;
  uint32_t *key_in = nullptr, *key_out = nullptr, *val_in = nullptr, *val_out = nullptr;
  void *tmp = nullptr;
  bool ok = cudaMalloc(reinterpret_cast<void **>(&pl->pairs), (size_t)E * sizeof(uint2)) == cudaSuccess &&
            cudaMalloc(reinterpret_cast<void **>(&pl->voff), ((size_t)n_keys + 1) * sizeof(uint32_t)) == cudaSuccess;
  const unsigned blocks = (unsigned)std::min<uint64_t>(((uint64_t)E + 255) / 256, (uint64_t)sm_count() * 32);
  if (ok && n_slabs == 1) {
    plan_pairs_kernel<<<blocks, 256, 0, st>>>(nullptr, indices, weight, slot_of, index_base, E, pl->pairs);
    count_launch();
    ok = cudaMemcpyAsync(pl->voff, offsets, ((size_t)n_rows + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st) ==
         cudaSuccess;
    pl->slab_edge[0] = 0;
    pl->slab_edge[1] = n_edges;
  } else if (ok) {
    size_t tmp_bytes = 0;
    int bits = 1;
    while (bits < 32 && (1ull << bits) < (uint64_t)n_keys)
      bits++;
    ok = cudaMalloc(reinterpret_cast<void **>(&key_in), (size_t)E * 4) == cudaSuccess &&
         cudaMalloc(reinterpret_cast<void **>(&key_out), (size_t)E * 4) == cudaSuccess &&
         cudaMalloc(reinterpret_cast<void **>(&val_in), (size_t)E * 4) == cudaSuccess &&
         cudaMalloc(reinterpret_cast<void **>(&val_out), (size_t)E * 4) == cudaSuccess &&
         cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key_in, key_out, val_in, val_out, (int64_t)E, 0, bits, st) ==
             cudaSuccess &&
         cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16) == cudaSuccess;
    if (ok) {
      plan_keys_kernel<<<blocks, 256, 0, st>>>(offsets, indices, slot_of, index_base, n_rows, E, pl->slab_rows,
                                               (uint32_t)n_slabs, key_in, val_in);
      count_launch();
      // LSD radix sort: stable, so edges of one (slab, row) segment keep their original order
      ok = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key_in, key_out, val_in, val_out, (int64_t)E, 0, bits, st) ==
           cudaSuccess;
    }
    if (ok) {
      plan_pairs_kernel<<<blocks, 256, 0, st>>>(val_out, indices, weight, slot_of, index_base, E, pl->pairs);
      count_launch();
      const unsigned kb = (unsigned)std::min<uint64_t>(((uint64_t)n_keys + 256) / 256, (uint64_t)sm_count() * 32);
      plan_offsets_kernel<<<kb, 256, 0, st>>>(key_out, E, n_keys, pl->voff);
      count_launch();
      std::vector<uint32_t> h(n_slabs + 1);
      ok = cudaStreamSynchronize(st) == cudaSuccess;
      for (int s = 0; s <= n_slabs && ok; s++)
        ok = cudaMemcpy(&h[s], pl->voff + (size_t)s * n_rows, 4, cudaMemcpyDeviceToHost) == cudaSuccess;
      for (int s = 0; s <= n_slabs; s++)
        pl->slab_edge[s] = h[s];
    }
  }
  if (ok)
    ok = cudaStreamSynchronize(st) == cudaSuccess && cudaGetLastError() == cudaSuccess;
  cudaFree(key_in), cudaFree(key_out), cudaFree(val_in), cudaFree(val_out), cudaFree(tmp);
  if (!ok) {
    fail(-1, "nts_gather_plan_create: device allocation or preprocessing failed", __FILE__, __LINE__);
    cudaFree(pl->pairs), cudaFree(pl->voff);
    delete pl;
    return nullptr;
  }
  return pl;
}

// Several chunks merged into ONE plan (the exchange engine aggregates all remote chunks of a rank in one launch when
// their rows arrive faster than one chunk computes): a stable sort of all parts' edges by (slab, output row); inside a
// segment the parts follow each other in the order given, each keeping its own edge order.
static nts_gather_plan *build_plan_parts(const nts_plan_part *parts, int n_parts, nts_vid_t n_rows, nts_vid_t gather_rows,
                                         int n_slabs, cudaStream_t st) {
  auto bad = [](const char *m) -> nts_gather_plan * {
    fail(-1, m, __FILE__, __LINE__);
    return nullptr;
  };
  uint64_t total = 0;
  for (int k = 0; k < n_parts; k++) {
    if (parts[k].n_edges && !(parts[k].offsets && parts[k].indices))
      return bad("null graph array in a plan part");
    if ((uint64_t)parts[k].row_add + parts[k].n_rows > n_rows)
      return bad("plan part rows exceed the output rows");
    total += parts[k].n_edges;
  }
  if (total >= 0xffffffffull)
    return bad("merged edge count must fit uint32 offsets");
  if (n_slabs < 1 || gather_rows == 0)
    n_slabs = 1;
  if ((uint64_t)n_slabs * n_rows >= 0xffffffffull)
    return bad("slabs * rows must fit 32-bit segment keys");
  nts_gather_plan *pl = new nts_gather_plan();
  pl->n_rows = n_rows;
  pl->n_edges = total;
  pl->gather_rows = gather_rows;
  pl->slabs = n_slabs;
  pl->slab_rows = gather_rows ? (gather_rows + n_slabs - 1) / n_slabs : 1;
  pl->slab_edge.assign(n_slabs + 1, 0);
  if (n_rows == 0 || total == 0)
    return pl;
  const uint32_t E = (uint32_t)total, n_keys = (uint32_t)n_slabs * n_rows;
  uint32_t *key_in = nullptr, *key_out = nullptr, *val_in = nullptr, *val_out = nullptr;
  uint2 *unsorted = nullptr;
  void *tmp = nullptr;
  size_t tmp_bytes = 0;
  int bits = 1;
  while (bits < 32 && (1ull << bits) < (uint64_t)n_keys)
    bits++;
  bool ok = cudaMalloc(reinterpret_cast<void **>(&pl->pairs), (size_t)E * sizeof(uint2)) == cudaSuccess &&
            cudaMalloc(reinterpret_cast<void **>(&pl->voff), ((size_t)n_keys + 1) * sizeof(uint32_t)) == cudaSuccess &&
            cudaMalloc(reinterpret_cast<void **>(&unsorted), (size_t)E * sizeof(uint2)) == cudaSuccess &&
            cudaMalloc(reinterpret_cast<void **>(&key_in), (size_t)E * 4) == cudaSuccess &&
            cudaMalloc(reinterpret_cast<void **>(&key_out), (size_t)E * 4) == cudaSuccess &&
            cudaMalloc(reinterpret_cast<void **>(&val_in), (size_t)E * 4) == cudaSuccess &&
            cudaMalloc(reinterpret_cast<void **>(&val_out), (size_t)E * 4) == cudaSuccess &&
            cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key_in, key_out, val_in, val_out, (int64_t)E, 0, bits, st) ==
                cudaSuccess &&
            cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16) == cudaSuccess;
  uint32_t e_off = 0;
  for (int k = 0; k < n_parts && ok; k++) {
    const nts_plan_part &pt = parts[k];
    if (!pt.n_edges)
      continue;
    const unsigned blocks = (unsigned)std::min<uint64_t>((pt.n_edges + 255) / 256, (uint64_t)sm_count() * 32);
    plan_part_keys_kernel<<<blocks, 256, 0, st>>>(pt.offsets, pt.indices, pt.slot_of, pt.weight, pt.index_base,
                                                  pt.index_add, pt.n_rows, pt.row_add, (uint32_t)pt.n_edges, e_off, n_rows,
                                                  pl->slab_rows, (uint32_t)n_slabs, key_in, val_in, unsorted);
    count_launch();
    e_off += (uint32_t)pt.n_edges;
  }
  if (ok)
    ok = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key_in, key_out, val_in, val_out, (int64_t)E, 0, bits, st) ==
         cudaSuccess;
  if (ok) {
    const unsigned blocks = (unsigned)std::min<uint64_t>(((uint64_t)E + 255) / 256, (uint64_t)sm_count() * 32);
    plan_permute_pairs_kernel<<<blocks, 256, 0, st>>>(val_out, unsorted, E, pl->pairs);
    count_launch();
    const unsigned kb = (unsigned)std::min<uint64_t>(((uint64_t)n_keys + 256) / 256, (uint64_t)sm_count() * 32);
    plan_offsets_kernel<<<kb, 256, 0, st>>>(key_out, E, n_keys, pl->voff);
    count_launch();
    std::vector<uint32_t> h(n_slabs + 1);
    ok = cudaStreamSynchronize(st) == cudaSuccess;
    for (int s = 0; s <= n_slabs && ok; s++)
      ok = cudaMemcpy(&h[s], pl->voff + (size_t)s * n_rows, 4, cudaMemcpyDeviceToHost) == cudaSuccess;
    for (int s = 0; s <= n_slabs; s++)
      pl->slab_edge[s] = h[s];
  }
  if (ok)
    ok = cudaStreamSynchronize(st) == cudaSuccess && cudaGetLastError() == cudaSuccess;
  cudaFree(key_in), cudaFree(key_out), cudaFree(val_in), cudaFree(val_out), cudaFree(tmp), cudaFree(unsorted);
  if (!ok) {
    fail(-1, "nts_gather_plan_create_parts: device allocation or preprocessing failed", __FILE__, __LINE__);
    cudaFree(pl->pairs), cudaFree(pl->voff);
    delete pl;
    return nullptr;
  }
  return pl;
}

// n_slabs >= 1: that slab count; n_slabs == 0: measured for feature_size like nts_gather_plan_create_tuned
nts_gather_plan *nts_gather_plan_create_parts(const nts_plan_part *parts, int n_parts, nts_vid_t n_rows,
                                              nts_vid_t gather_rows, int n_slabs, nts_vid_t feature_size, void *stream) {
  if (!parts || n_parts < 1) {
    fail(-1, "no plan parts", __FILE__, __LINE__);
    return nullptr;
  }
  cudaStream_t st = as_stream(stream);
  if (n_slabs >= 1)
    return build_plan_parts(parts, n_parts, n_rows, gather_rows, n_slabs, st);
  uint64_t total = 0;
  for (int k = 0; k < n_parts; k++)
    total += parts[k].n_edges;
  const int s_max = nts_gather_plan_pick_slabs(gather_rows, total, n_rows, feature_size, 16ull << 20);
  nts_gather_plan *best = build_plan_parts(parts, n_parts, n_rows, gather_rows, 1, st);
  if (!best || s_max <= 1 || total == 0 || feature_size == 0)
    return best;
  float *x = nullptr, *y = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const size_t xb = (size_t)gather_rows * feature_size * sizeof(float), yb = (size_t)n_rows * feature_size * sizeof(float);
  bool ok = cudaMalloc(reinterpret_cast<void **>(&x), xb) == cudaSuccess &&
            cudaMalloc(reinterpret_cast<void **>(&y), yb) == cudaSuccess &&
            cudaMemsetAsync(x, 0, xb, st) == cudaSuccess && cudaMemsetAsync(y, 0, yb, st) == cudaSuccess &&
            cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess;
  auto time_plan = [&](nts_gather_plan *pl, float *ms) -> bool {
    *ms = 1e30f;
    for (int it = 0; it < 3; it++) {
      float t = 0.f;
      if (cudaEventRecord(e0, st) != cudaSuccess || run_plan(pl, x, y, feature_size, st) != 0 ||
          cudaEventRecord(e1, st) != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess ||
          cudaEventElapsedTime(&t, e0, e1) != cudaSuccess)
        return false;
      if (it > 0 && t < *ms)
        *ms = t;
    }
    return true;
  };
  float best_ms = 0.f;
  ok = ok && time_plan(best, &best_ms);
  for (int s = 2; ok; s *= 2) {
    const int cand = s > s_max ? s_max : s;
    nts_gather_plan *pl = build_plan_parts(parts, n_parts, n_rows, gather_rows, cand, st);
    float ms = 0.f;
    if (!pl || !time_plan(pl, &ms)) {
      nts_gather_plan_destroy(pl);
      break;
    }
    if (ms < best_ms) {
      nts_gather_plan_destroy(best);
      best = pl;
      best_ms = ms;
    } else {
      nts_gather_plan_destroy(pl);
      if (ms > 1.1f * best_ms)
        break;
    }
    if (cand == s_max)
      break;
  }
  cudaFree(x), cudaFree(y);
  if (e0)
    cudaEventDestroy(e0);
  if (e1)
    cudaEventDestroy(e1);
  best->tuned_ms = best_ms;
  return best;
}

float nts_gather_plan_tuned_ms(const nts_gather_plan *pl) { return pl ? pl->tuned_ms : 0.f; }

// Slab count by measurement.  Whether bucketing pays depends on how skewed the gathered rows are (hub sources stay in
// L1/L2 by themselves: on the Zipf graph of config B one launch at F=602 takes 12.8 ms unbucketed, 20.9 ms with 14
// slabs; with uniform endpoints 35.3 ms vs 16.9 ms) and on the degree distribution of the output rows (every non-empty
// (slab, row) segment costs a read-modify-write of the output row) - so the candidates 1, 2, 4, ... up to the
// size-based bound are built and timed on the real arrays (zero features: the access pattern does not depend on the
// values), 1 warm + 2 timed launches each, and the fastest is kept.  One-time cost per (chunk, direction, width).
nts_gather_plan *nts_gather_plan_create_tuned(const nts_vid_t *offsets, const nts_vid_t *indices, const float *weight,
                                              const nts_vid_t *slot_of, nts_vid_t index_base, nts_vid_t n_rows,
                                              uint64_t n_edges, nts_vid_t gather_rows, nts_vid_t feature_size,
                                              void *stream) {
  cudaStream_t st = as_stream(stream);
  const int s_max = nts_gather_plan_pick_slabs(gather_rows, n_edges, n_rows, feature_size, 16ull << 20);
  nts_gather_plan *best = nts_gather_plan_create(offsets, indices, weight, slot_of, index_base, n_rows, n_edges,
                                                 gather_rows, 1, stream);
  if (!best || s_max <= 1 || n_edges == 0 || feature_size == 0)
    return best;
  float *x = nullptr, *y = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const size_t xb = (size_t)gather_rows * feature_size * sizeof(float), yb = (size_t)n_rows * feature_size * sizeof(float);
  bool ok = cudaMalloc(reinterpret_cast<void **>(&x), xb) == cudaSuccess &&
            cudaMalloc(reinterpret_cast<void **>(&y), yb) == cudaSuccess &&
            cudaMemsetAsync(x, 0, xb, st) == cudaSuccess && cudaMemsetAsync(y, 0, yb, st) == cudaSuccess &&
            cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess;
  auto time_plan = [&](nts_gather_plan *pl, float *ms) -> bool {
    *ms = 1e30f;
    for (int it = 0; it < 3; it++) {
      float t = 0.f;
      if (cudaEventRecord(e0, st) != cudaSuccess || run_plan(pl, x, y, feature_size, st) != 0 ||
          cudaEventRecord(e1, st) != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess ||
          cudaEventElapsedTime(&t, e0, e1) != cudaSuccess)
        return false;
      if (it > 0 && t < *ms)
        *ms = t;
    }
    return true;
  };
  float best_ms = 0.f;
  ok = ok && time_plan(best, &best_ms);
  for (int s = 2; ok; s *= 2) {
    const int cand = s > s_max ? s_max : s;
    nts_gather_plan *pl = nts_gather_plan_create(offsets, indices, weight, slot_of, index_base, n_rows, n_edges,
                                                 gather_rows, cand, stream);
    float ms = 0.f;
    if (!pl || !time_plan(pl, &ms)) {
      nts_gather_plan_destroy(pl);
      break;
    }
    if (ms < best_ms) {
      nts_gather_plan_destroy(best);
      best = pl;
      best_ms = ms;
    } else {
      nts_gather_plan_destroy(pl);
      if (ms > 1.1f * best_ms) // getting worse: larger slab counts only add read-modify-writes
        break;
    }
    if (cand == s_max)
      break;
  }
  cudaFree(x), cudaFree(y);
  if (e0)
    cudaEventDestroy(e0);
  if (e1)
    cudaEventDestroy(e1);
  best->tuned_ms = best_ms;
  return best;
}

int nts_gather_plan_destroy(nts_gather_plan *pl) {
  if (!pl)
    return 0;
  cudaFree(pl->pairs);
  cudaFree(pl->voff);
  cudaFree(pl->workspace);
  delete pl;
  return 0;
}

int nts_gather_plan_slabs(const nts_gather_plan *pl) { return pl ? pl->slabs : 0; }

uint64_t nts_gather_plan_bytes(const nts_gather_plan *pl) {
  if (!pl)
    return 0;
  return pl->n_edges * 8ull + ((uint64_t)pl->slabs * pl->n_rows + 1) * 4ull + pl->workspace_floats * 4ull;
}

int nts_gather_plan_last_launch(const nts_gather_plan *pl, int *launches, int *grid, int *k, int *u, int *outv) {
  NTS_ARG_CHECK(pl != nullptr, "null plan");
  if (launches)
    *launches = pl->last_launches;
  if (grid)
    *grid = pl->last_grid;
  if (k)
    *k = pl->last_k;
  if (u)
    *u = pl->last_u;
  if (outv)
    *outv = pl->last_outv;
  return 0;
}

int nts_gather_plan_run(nts_gather_plan *pl, const float *input, float *output, nts_vid_t feature_size, void *stream) {
  NTS_ARG_CHECK(pl != nullptr, "null plan");
  return run_plan(pl, input, output, feature_size, as_stream(stream));
}

int nts_gather_plan_set_variant(int variant) {
  NTS_ARG_CHECK(variant == 0 || variant == 1, "variant must be 0 (register staging) or 1 (TMA row staging)");
  g_plan_variant = variant;
  return 0;
}

int nts_gather_plan_set_tuning(int u, int min_blocks, int edges_per_warp) {
  NTS_ARG_CHECK(u >= 0 && min_blocks >= 0 && edges_per_warp >= 0 && edges_per_warp <= 4096, "bad tuning point");
  g_plan_u = u;
  g_plan_minb = min_blocks;
  g_plan_q = edges_per_warp;
  return 0;
}

} // extern "C"
