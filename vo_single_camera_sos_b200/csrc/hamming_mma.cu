// Tensor-core engine of the brute-force Hamming matcher (step 2 of the SOS front-end; same contract as the XOR+POPC
// engine in hamming.cu, which it replaces for large problems: FeatureMatcher.match on cv2.BFMatcher(NORM_HAMMING),
// reference omnistereo/camera_models.py:402-446).
//
// Hamming distance is a contraction: with the 256 descriptor bits mapped to signs s = +-1,
//     <s_q, s_t> = 256 - 2 d(q, t),
// so the all-pairs distance matrix of a (query, train) segment is an integer GEMM with K = 256.  Here it runs on the
// 5th-generation tensor cores (tcgen05.mma kind::i8, int32 accumulators in tensor memory: exact), and the part that bounds
// the POPC engine — ~26 integer instructions per descriptor pair — shrinks to reading the accumulator back and ONE max:
//
//   * operands are +-16 instead of +-1, so the accumulator holds 256 * <s_q, s_t>;
//   * 32 extra K columns carry bookkeeping through the same MMA: the query side holds the constants (16, 1, 127 x 30),
//     the train side (-(u >> 4), -(u & 15), pad x 30) with u = row of the train descriptor inside its 128-row tile and
//     pad = -128 for the padding rows of a ragged last tile, 0 otherwise.  The accumulator therefore is
//         acc = 256 * <s_q, s_t> - u - 487680 * [padding row]
//     i.e. ALREADY the packed sort key: max(acc) over a tile = smallest distance, ties to the lowest train row (OpenCV's
//     order), padding rows lose against everything — no index arithmetic, no masking in the epilogue;
//   * the epilogue thread that owns query row r (TMEM lane r) reads its 128 accumulators with tcgen05.ld and keeps the
//     running maximum (3-input integer max: half an instruction per pair); top-2 costs 3 instructions per pair.
//
// Data flow per work item (256 query rows x a range of train tiles; one CTA per SM at a time, warp-specialised):
//   expand kernel : descriptors -> "tile-ready" int8 images in global memory, 128 rows x 288 bytes per tile, stored in the
//                   canonical K-major no-swizzle UMMA layout (8-row x 16-byte core matrices), so that a tile is ONE
//                   contiguous 36 KB block
//   warp 0        : cp.async.bulk (TMA engine, mbarrier complete_tx) of the two query tiles, then a 4-stage ring of
//                   train tiles
//   warps 1, 3    : one lane of each issues 9 tcgen05.mma (M128 N128 K32) per train tile, one warp per query tile, into a double-buffered TMEM
//                   accumulator (2 buffers x 2 query tiles x 128 columns = all 512 columns); tcgen05.commit releases the
//                   shared-memory stage and publishes the accumulator
//   warp 2        : tensor-memory allocation
//   warps 4..11   : epilogue, one thread per (query tile, row)
// The per-split partial keys use the POPC engine's format and are merged by the same hamming_merge_kernel.
#include <cuda_bf16.h>

#include "sos_common.cuh"
#include "tc_common.cuh"

namespace sos_hamming_mma {

constexpr int TILE = 128;                        // rows per expanded tile = UMMA M = UMMA N
constexpr int KB = 288;                          // bytes of K per row: 256 sign bytes + 32 bookkeeping bytes
constexpr int CHUNKS = KB / 16;                  // 18 core-matrix columns
constexpr int GROUP_BYTES = CHUNKS * 128;        // one 8-row group: 18 core matrices of 8 x 16 bytes
constexpr int TILE_BYTES = (TILE / 8) * GROUP_BYTES;  // 36864
constexpr int THREADS = 384;
constexpr int smem_bytes(int stages) { return (2 + stages) * TILE_BYTES + 256; }
constexpr int PAD_DROP = 30 * 127 * 128;         // what the 30 pad columns subtract from a padding row's accumulator
constexpr int KEY_IDX_BITS = 22;                 // must match hamming.cu
constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;
constexpr int ITEM_SPLIT_BITS = 6, ITEM_TILE_BITS = 10;   // item = seg << 16 | tile << 6 | split (hamming.cu)

// ---- expansion ------------------------------------------------------------------------------------------------------
// four descriptor bits -> four int8 values: bit 0 -> +16, bit 1 -> -16
__device__ __forceinline__ uint32_t expand4(uint32_t nib) {
  const uint32_t spread = (nib * 0x00204081u) & 0x01010101u;   // bit i -> LSB of byte i (the shifted copies never overlap)
  return 0x10101010u ^ (spread * 0xE0u);
}

// role 0: query rows (A operand), role 1: train rows (B operand).  Tile t of segment s lives at
// out + (s * tiles_per_seg + t) * TILE_BYTES; byte (row r, k) of a tile at (r / 8) * GROUP_BYTES + (k / 16) * 128 + (r % 8) * 16 + k % 16.
// One launch expands both roles (blockIdx.z); the nine 16-byte units of a thread are independent, so the loop is unrolled
// and all descriptor loads are in flight before the first store (the rolled loop was latency bound: 4 x 27 us per step).
struct ExpandArgs {
  const uint32_t* desc[2];
  const int32_t *start[2], *len[2];
  int max_rows[2], tiles_per_seg[2];
  uint8_t* out[2];
};

__global__ void __launch_bounds__(256)
expand_kernel(const ExpandArgs a) {
  const int role = blockIdx.z, seg = blockIdx.y, tile = blockIdx.x;
  if (tile >= a.tiles_per_seg[role]) return;
  const int n = min(a.len[role][seg], a.max_rows[role]);
  const int row0 = tile * TILE;
  if (row0 >= n) return;
  const uint32_t* d = a.desc[role] + (size_t)a.start[role][seg] * 8;
  uint8_t* tbase = a.out[role] + ((size_t)seg * a.tiles_per_seg[role] + tile) * TILE_BYTES;
  constexpr int UNITS = TILE * CHUNKS / 256;   // 9
  uint32_t bits[UNITS];
#pragma unroll
  for (int it = 0; it < UNITS; ++it) {
    const int u = threadIdx.x + it * 256;
    const int r8 = u & 7, chunk = (u >> 3) % CHUNKS, rg = u / (8 * CHUNKS);
    const int row = row0 + rg * 8 + r8;
    bits[it] = (chunk < 16 && row < n) ? (__ldg(d + (size_t)row * 8 + (chunk >> 1)) >> ((chunk & 1) * 16)) & 0xFFFFu : 0u;
  }
#pragma unroll
  for (int it = 0; it < UNITS; ++it) {
    const int u = threadIdx.x + it * 256;
    const int r8 = u & 7, chunk = (u >> 3) % CHUNKS, rg = u / (8 * CHUNKS);
    const int r = rg * 8 + r8, row = row0 + r;
    const bool valid = row < n;
    uint4 v;
    if (chunk < 16) {
      v.x = expand4(bits[it] & 15u);
      v.y = expand4((bits[it] >> 4) & 15u);
      v.z = expand4((bits[it] >> 8) & 15u);
      v.w = expand4(bits[it] >> 12);
      if (!valid) v = make_uint4(0u, 0u, 0u, 0u);          // padding rows: <s_q, s_t> = 0
    } else if (role == 0) {
      v = make_uint4(0x7F7F7F7Fu, 0x7F7F7F7Fu, 0x7F7F7F7Fu, 0x7F7F7F7Fu);
      if (chunk == 16) v.x = 0x7F7F0110u;                  // bytes (16, 1, 127, 127)
    } else {
      const uint32_t pad = valid ? 0u : 0x80808080u;       // -128 in every pad column of a padding row
      v = make_uint4(pad, pad, pad, pad);
      if (chunk == 16) {
        const uint32_t hi = (uint32_t)(-(r >> 4)) & 0xFFu, lo = (uint32_t)(-(r & 15)) & 0xFFu;
        v.x = (pad & 0xFFFF0000u) | (lo << 8) | hi;
      }
    }
    *(uint4*)(tbase + (size_t)rg * GROUP_BYTES + chunk * 128 + r8 * 16) = v;
  }
}

using namespace sos_tc;
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) { return sos_tc::smem_desc(addr, (uint32_t)GROUP_BYTES); }
// instruction descriptor: D = S32 (2 << 4), A = B = signed int8 (1 << 7, 1 << 10), both K-major, N = 128, M = 128
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TILE >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
// the float-descriptor L2 engine: D = F32 (1 << 4), A = B = BF16 (1 << 7, 1 << 10), both K-major, N = 128, M = 128
constexpr uint32_t IDESC_L2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TILE >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);

template <bool TOP2>
__device__ __forceinline__ void reduce32(const int (&v)[32], int& m0, int& m1) {
  if (TOP2) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      m1 = max(m1, min(m0, v[i]));
      m0 = max(m0, v[i]);
    }
  } else {
    int a = m0;
#pragma unroll
    for (int i = 0; i < 32; i += 2) a = __vimax3_s32(a, v[i], v[i + 1]);
    m0 = a;
  }
}

// accumulator of (query row, train row u of tile `tile`) -> the POPC engine's key (distance << 22 | train row of the segment)
__device__ __forceinline__ uint32_t acc_to_key(int acc, int tile) {
  const int dot = (acc + 255) >> 8;              // acc = 256 * dot - u with 0 <= u < 128
  if (dot < -256) return KEY_NONE;               // padding row (or no train row at all)
  const int u = 256 * dot - acc;
  return ((uint32_t)((256 - dot) >> 1) << KEY_IDX_BITS) | (uint32_t)(tile * TILE + u);
}

// ---- float-descriptor L2 engine: epilogue pieces ------------------------------------------------------------------------
// accumulator (float32, an exact integer) = 2 <q, t> - |t|^2 of (query row, train row u); the largest one is the nearest
// neighbour.  key = acc * 128 + (127 - u): max(key) = largest accumulator, ties to the lowest train row.
constexpr int L2_PAD_BELOW = -12000000;   // valid accumulators are >= -128 * 255^2 = -8323200; padding rows sit at -16.7 M
template <bool TOP2>
__device__ __forceinline__ void reduce32_l2(const int (&v)[32], int col0, int& m0, int& m1) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int k = __float2int_rn(__int_as_float(v[i])) * 128 + (127 - (col0 + i));
    if (TOP2) m1 = max(m1, min(m0, k));
    m0 = max(m0, k);
  }
}
constexpr unsigned long long KEY64_NONE = ~0ull;
// key of the tile -> (squared distance << 32 | train row of the segment); nq = |q|^2 of this thread's query row
__device__ __forceinline__ unsigned long long l2_key(int key, int tile, int nq) {
  const int acc = key >> 7;                          // floor: key = acc * 128 + low, 0 <= low < 128
  if (acc < L2_PAD_BELOW) return KEY64_NONE;
  const int u = 127 - (key & 127);
  return ((unsigned long long)(uint32_t)(nq - acc) << 32) | (unsigned long long)(uint32_t)(tile * TILE + u);
}

struct Args {
  const int32_t* a_norm;   // KIND_L2: |q|^2 per expanded query row [n_seg, q_tiles_per_seg * 128]
  const uint8_t *a_exp, *b_exp;
  const int32_t *q_len, *t_len;
  int max_nq, max_nt, splits, q_tiles_per_seg, t_tiles_per_seg;
  const uint32_t* items;
  const int32_t* n_items;
  void* partial;           // uint2 (Hamming keys) or ulonglong2 (L2 keys) per (segment, query row, split)
};

// Work item -> (segment, first query row, split, train tiles [tile_begin, tile_begin + n_iter))
struct Item {
  int seg, q_tile, split, nq, tile_begin, n_iter;
};
__device__ __forceinline__ Item decode_item(const Args& a, uint32_t item) {
  Item it;
  it.seg = (int)(item >> 16);
  it.q_tile = (int)((item >> ITEM_SPLIT_BITS) & ((1u << ITEM_TILE_BITS) - 1u));   // 256 query rows
  it.split = (int)(item & ((1u << ITEM_SPLIT_BITS) - 1u));
  it.nq = min(a.q_len[it.seg], a.max_nq);
  const int nt = min(a.t_len[it.seg], a.max_nt);
  const int t_tiles = (nt + TILE - 1) / TILE;
  const int chunk_tiles = (t_tiles + a.splits - 1) / a.splits;
  it.tile_begin = min(t_tiles, it.split * chunk_tiles);
  it.n_iter = min(t_tiles, it.tile_begin + chunk_tiles) - it.tile_begin;
  return it;
}

// A CTA walks the work list with stride gridDim.x (launched either with one CTA per item, or PERSISTENT with one CTA per SM:
// see the host side).  Barriers, the tensor-memory allocation and the three pipelines (train-tile ring, accumulator double
// buffer, query tiles) live across items, so in the persistent shape the tensor pipe only idles while the next item's query
// tiles land (the train tiles of the next item are already streaming into the ring by then).  The roles agree on every
// counter because they walk the same items in the same order:
//   t  = train tiles consumed so far by this CTA -> ring stage t % STAGES, accumulator buffer t & 1
//   ic = items with work so far                  -> phase of the query-tile barriers
template <bool TOP2, int STAGES, int KIND>
__global__ void __launch_bounds__(THREADS, 1) mma_kernel(Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int n_items = *a.n_items;
  if ((int)blockIdx.x >= n_items) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 2 * TILE_BYTES;
  uint64_t* bars = (uint64_t*)(smem + (2 + STAGES) * TILE_BYTES);
  // bars: FULL + s (train stage landed), EMPTY + s (stage consumed), TFULL + b (accumulator complete), TEMPTY + b
  // (accumulator drained), AFULL (query tiles landed), AEMPTY (query tiles no longer read by the tensor core)
  constexpr int FULL = 0, EMPTY = STAGES, TFULL = 2 * STAGES, TEMPTY = 2 * STAGES + 2, AFULL = 2 * STAGES + 4, AEMPTY = 2 * STAGES + 5;
  uint32_t* tmem_slot = (uint32_t*)(bars + AEMPTY + 1);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(BAR(FULL + i), 1); mbar_init(BAR(EMPTY + i), 2); }   // EMPTY, TFULL, AEMPTY: one commit per MMA warp
    mbar_init(BAR(TFULL), 2); mbar_init(BAR(TFULL + 1), 2);
    mbar_init(BAR(TEMPTY), 8); mbar_init(BAR(TEMPTY + 1), 8);
    mbar_init(BAR(AFULL), 1); mbar_init(BAR(AEMPTY), 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // producer: the whole (converged) warp walks the items, the elected lane issues the copies (tc_common.cuh: elect_one)
    uint32_t t = 0, ic = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const Item it = decode_item(a, a.items[w]);
      if (it.n_iter <= 0) continue;
      const uint8_t* gb = a.b_exp + ((size_t)it.seg * a.t_tiles_per_seg + it.tile_begin) * TILE_BYTES;
      auto load_b = [&](int k) {
        const uint32_t s = t % STAGES;
        mbar_wait(BAR(EMPTY + s), ((t / STAGES) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(BAR(FULL + s), TILE_BYTES);
          const uint32_t dst = smem_u32(sB) + s * TILE_BYTES;
#pragma unroll
          for (int p = 0; p < 4; ++p)
            bulk_g2s(dst + p * (TILE_BYTES / 4), gb + (size_t)k * TILE_BYTES + (size_t)p * (TILE_BYTES / 4), TILE_BYTES / 4, BAR(FULL + s));
        }
        __syncwarp();
        ++t;
      };
      // the first train tiles of this item go into the ring as soon as stages free up, i.e. while the previous item is still
      // being multiplied; only then wait for the tensor core to be done with the previous item's query tiles and fetch ours
      const int pre = min(it.n_iter, STAGES);
      for (int k = 0; k < pre; ++k) load_b(k);
      mbar_wait(BAR(AEMPTY), (ic & 1) ^ 1);
      const uint8_t* ga = a.a_exp + ((size_t)it.seg * a.q_tiles_per_seg + (size_t)it.q_tile * 2) * TILE_BYTES;
      if (elect_one()) {
        mbar_expect_tx(BAR(AFULL), 2u * TILE_BYTES);
#pragma unroll
        for (int p = 0; p < 8; ++p) bulk_g2s(smem_u32(sA) + p * (TILE_BYTES / 4), ga + (size_t)p * (TILE_BYTES / 4), TILE_BYTES / 4, BAR(AFULL));
      }
      __syncwarp();
      ++ic;
      for (int k = pre; k < it.n_iter; ++k) load_b(k);
    }
  } else if (warp == 1 || warp == 3) {
    // MMA issuers: converged warps, elected lane — issued from `if (lane == 0)` every tcgen05.mma was wrapped in an
    // elect / broadcast loop of ~13 instructions and the issuing thread, not the tensor pipe, set the pace.  Two of them (on
    // different sub-partitions), one per query tile = accumulator: the issue loop of a single warp sharing its scheduler
    // with busy epilogue warps was still slower than the tensor pipe (found on the score engine, score_mma.cuh).
    const int qt = warp == 1 ? 0 : 1;
    const uint64_t da = smem_desc(smem_u32(sA) + qt * TILE_BYTES);
    uint32_t t = 0, ic = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const Item it = decode_item(a, a.items[w]);
      if (it.n_iter <= 0) continue;
      mbar_wait(BAR(AFULL), ic & 1);
      ++ic;
      for (int k = 0; k < it.n_iter; ++k, ++t) {
        const uint32_t s = t % STAGES, b = t & 1;
        mbar_wait(BAR(TEMPTY + b), ((t >> 1) & 1) ^ 1);    // the epilogue has drained this accumulator buffer
        mbar_wait(BAR(FULL + s), (t / STAGES) & 1);        // the train tile has landed
        tc_fence_after();
        const uint64_t db = smem_desc(smem_u32(sB) + s * TILE_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < KB / 32; ++kk)
            tc_mma<KIND>(tmem + (b * 2 + qt) * TILE, da + (uint64_t)(16 * kk), db + (uint64_t)(16 * kk), KIND == KIND_L2 ? IDESC_L2 : IDESC, kk > 0);
          tc_commit(BAR(EMPTY + s));    // shared-memory stage free once both warps' MMAs have read it
          tc_commit(BAR(TFULL + b));    // this query tile's accumulator complete
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(BAR(AEMPTY));         // every MMA of this item has read the query tiles
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int g = (warp - 4) >> 2, quarter = warp & 3;   // a warp may only touch the TMEM lanes of its quarter
    uint32_t t = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const Item it = decode_item(a, a.items[w]);
      const int row = it.q_tile * 256 + g * TILE + quarter * 32 + lane;
      uint32_t k0 = KEY_NONE, k1 = KEY_NONE;
      unsigned long long w0 = KEY64_NONE, w1 = KEY64_NONE;
      int nq_norm = 0;
      if (KIND == KIND_L2 && it.n_iter > 0) nq_norm = a.a_norm[(size_t)it.seg * a.q_tiles_per_seg * TILE + row];
      for (int k = 0; k < it.n_iter; ++k, ++t) {
        const uint32_t b = t & 1;
        mbar_wait(BAR(TFULL + b), (t >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + (b * 2 + g) * TILE;
        int m0 = INT_MIN, m1 = INT_MIN;
        int va[32], vb[32];
        auto reduce = [&](const int (&v)[32], int col0) {
          if (KIND == KIND_L2) reduce32_l2<TOP2>(v, col0, m0, m1);
          else reduce32<TOP2>(v, m0, m1);
        };
        tmem_ld32(taddr, va);
        tmem_ld_wait(va);
        tmem_ld32(taddr + 32, vb);          // in flight while va is reduced
        reduce(va, 0);
        tmem_ld_wait(vb);
        tmem_ld32(taddr + 64, va);
        reduce(vb, 32);
        tmem_ld_wait(va);
        tmem_ld32(taddr + 96, vb);
        reduce(va, 64);
        tmem_ld_wait(vb);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(TEMPTY + b));     // buffer b may be overwritten
        reduce(vb, 96);
        const int tile = it.tile_begin + k;
        if (KIND == KIND_L2) {
          const unsigned long long key0 = l2_key(m0, tile, nq_norm);
          if (TOP2) {
            const unsigned long long key1 = m1 == INT_MIN ? KEY64_NONE : l2_key(m1, tile, nq_norm);
            w1 = min(w1, max(w0, key0));
            w0 = min(w0, key0);
            w1 = min(w1, key1);
          } else {
            w0 = min(w0, key0);
          }
        } else {
          const uint32_t key0 = acc_to_key(m0, tile);
          if (TOP2) {
            const uint32_t key1 = acc_to_key(m1, tile);
            k1 = min(k1, max(k0, key0));                    // merge two sorted pairs
            k0 = min(k0, key0);
            k1 = min(k1, key1);
          } else {
            k0 = min(k0, key0);
          }
        }
      }
      // (no train rows in this split: the keys stay "none")
      if (row < it.nq) {
        const size_t o = ((size_t)it.seg * a.max_nq + row) * a.splits + it.split;
        if (KIND == KIND_L2) ((ulonglong2*)a.partial)[o] = make_ulonglong2(w0, w1);
        else ((uint2*)a.partial)[o] = make_uint2(k0, k1);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ---- float-descriptor L2 engine: expansion ---------------------------------------------------------------------------
// Float descriptors whose values are integers in [0, 255] (what cv2's SIFT returns) are exactly representable in bfloat16,
// their dot products (< 2^24) exactly in the float32 accumulator: |q - t|^2 = |q|^2 + |t|^2 - 2 <q, t> comes out of the
// tensor core as an exact integer.  A tile has the Hamming engine's geometry: 144 bf16 = 288 bytes per row; elements 0..127
// hold 2 q (query role) or t (train role), elements 128..130 carry the norm: (1, 256, 65536) against minus the base-256
// digits of |t|^2, so that the accumulator is 2 <q, t> - |t|^2; padding rows get the digits (255, 255, 255).
// One block per tile; every warp first sums the squares of 16 rows.  flag[0] is set when a value is not an integer in [0, 255].
__global__ void __launch_bounds__(256)
expand_l2_kernel(const float* __restrict__ desc, int dim, const int32_t* __restrict__ start, const int32_t* __restrict__ len,
                 int max_rows, int tiles_per_seg, int role, uint8_t* __restrict__ out, int32_t* __restrict__ norms,
                 int32_t* __restrict__ flag) {
  __shared__ int snorm[TILE];
  const int seg = blockIdx.y, tile = blockIdx.x;
  const int n = min(len[seg], max_rows);
  const int row0 = tile * TILE;
  if (row0 >= n) return;
  const float* d = desc + (size_t)start[seg] * dim;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  bool bad = false;
  for (int r = warp; r < TILE; r += 8) {
    const int row = row0 + r;
    int acc = 0;
    if (row < n)
      for (int k = lane; k < dim; k += 32) {
        const float v = __ldg(d + (size_t)row * dim + k);
        const int iv = (int)v;
        bad |= !(v >= 0.f && v <= 255.f && (float)iv == v);
        acc += iv * iv;
      }
    acc = __reduce_add_sync(0xFFFFFFFFu, acc);
    if (lane == 0) snorm[r] = row < n ? acc : -1;
  }
  if (bad) atomicOr(flag, 1);
  __syncthreads();
  uint8_t* tbase = out + ((size_t)seg * tiles_per_seg + tile) * TILE_BYTES;
  if (role == 0 && norms)
    for (int r = threadIdx.x; r < TILE; r += 256) norms[((size_t)seg * tiles_per_seg + tile) * TILE + r] = max(snorm[r], 0);
  const float scale = role == 0 ? 2.f : 1.f;
  for (int u = threadIdx.x; u < TILE * CHUNKS; u += 256) {
    const int r8 = u & 7, chunk = (u >> 3) % CHUNKS, rg = u / (8 * CHUNKS);
    const int r = rg * 8 + r8, row = row0 + r;
    const bool valid = row < n;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (chunk < 16) {                       // elements 8 chunk .. 8 chunk + 7
      if (valid) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int k = chunk * 8 + e;
          const float v = k < dim ? scale * __ldg(d + (size_t)row * dim + k) : 0.f;
          w[e >> 1] |= (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v)) << (16 * (e & 1));
        }
      }
    } else if (chunk == 16) {               // elements 128 .. 135: the norm digits
      float e0, e1, e2;
      if (role == 0) { e0 = 1.f; e1 = 256.f; e2 = 65536.f; }
      else {
        const int nt = valid ? snorm[r] : 0xFFFFFF;
        e0 = -(float)(nt & 255); e1 = -(float)((nt >> 8) & 255); e2 = -(float)((nt >> 16) & 255);
      }
      w[0] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(e0)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(e1)) << 16);
      w[1] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(e2));
    }
    *(uint4*)(tbase + (size_t)rg * GROUP_BYTES + chunk * 128 + r8 * 16) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

}  // namespace sos_hamming_mma

// Host side, called by sos_hamming_top2 (hamming.cu).  `ws` holds the partial keys + work items (already planned);
// `exp` is scratch for the expanded tiles.
size_t sos_hamming_mma_scratch_bytes(int n_seg, int max_nq, int max_nt) {
  using namespace sos_hamming_mma;
  const size_t qt = (size_t)((max_nq + 255) / 256) * 2, tt = (size_t)((max_nt + TILE - 1) / TILE);
  return (size_t)n_seg * (qt + tt) * TILE_BYTES + 256;
}

int sos_hamming_mma_launch(sos_ctx* ctx, const uint32_t* q, const uint32_t* t, const int32_t* q_start,
                           const int32_t* q_len, const int32_t* t_start, const int32_t* t_len, int n_seg, int max_nq,
                           int max_nt, int splits, const uint32_t* items, const int32_t* n_items, unsigned max_items,
                           bool top2, void* exp, uint2* partial) {
  using namespace sos_hamming_mma;
  const int qt = ((max_nq + 255) / 256) * 2, tt = (max_nt + TILE - 1) / TILE;
  uint8_t* a_exp = (uint8_t*)exp;
  uint8_t* b_exp = a_exp + (size_t)n_seg * qt * TILE_BYTES;
  if (qt > 0 || tt > 0) {
    ExpandArgs e;
    e.desc[0] = q; e.desc[1] = t; e.start[0] = q_start; e.start[1] = t_start; e.len[0] = q_len; e.len[1] = t_len;
    e.max_rows[0] = max_nq; e.max_rows[1] = max_nt; e.tiles_per_seg[0] = qt; e.tiles_per_seg[1] = tt;
    e.out[0] = a_exp; e.out[1] = b_exp;
    expand_kernel<<<dim3(qt > tt ? qt : tt, n_seg, 2), 256, 0, ctx->stream>>>(e);
    SOS_LAUNCHED_AS(ctx, "hamming_expand_kernel");
  }
  Args a;
  a.a_exp = a_exp; a.b_exp = b_exp; a.q_len = q_len; a.t_len = t_len;
  a.max_nq = max_nq; a.max_nt = max_nt; a.splits = splits; a.q_tiles_per_seg = qt; a.t_tiles_per_seg = tt;
  a.items = items; a.n_items = n_items; a.partial = partial;
  a.a_norm = nullptr;
  // ring depth 4 (221 KB of shared memory); 3 stages measured the same (profiles/r02/score_variants_and_ring_depth_ab.log)
  constexpr int ST = 4;
  static bool attr_set[64][2] = {};     // per device: the opt-in to > 48 KB of dynamic shared memory
  const int dev = ctx->device & 63, v = top2 ? 1 : 0;
  if (!attr_set[dev][v]) {
    if (top2) SOS_CUDA(cudaFuncSetAttribute(mma_kernel<true, ST, KIND_HAMMING>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(ST)));
    else SOS_CUDA(cudaFuncSetAttribute(mma_kernel<false, ST, KIND_HAMMING>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(ST)));
    attr_set[dev][v] = true;
  }
  // One kernel, two launch shapes.  Short items (stereo buckets: a handful of train tiles each) run PERSISTENT — one CTA per
  // SM walks the work list, so barriers, tensor memory and the pipelines survive from item to item (measured: 0.083 ->
  // 0.076 ms on the 384 C2 buckets).  Long items (temporal matching: ~36 train tiles each) get one CTA per item and leave
  // the balancing to the hardware scheduler: the fixed striding of the persistent shape loses more to unequal SMs than the
  // per-item set-up costs (0.285 vs 0.300 ms; profiles/r02/hamming_ab_persistent.log).
  const bool persistent = max_nt <= 2048;
  const unsigned grid = (persistent && max_items > (unsigned)ctx->sm_count) ? (unsigned)ctx->sm_count : max_items;
  if (top2) mma_kernel<true, ST, KIND_HAMMING><<<grid, THREADS, smem_bytes(ST), ctx->stream>>>(a);
  else mma_kernel<false, ST, KIND_HAMMING><<<grid, THREADS, smem_bytes(ST), ctx->stream>>>(a);
  SOS_LAUNCHED_AS(ctx, "hamming_mma_kernel");
  return SOS_OK;
}

// The float-descriptor L2 engine (sos_l2_top2 in hamming.cu).  `exp` = expanded tiles followed by the query norms.
size_t sos_l2_mma_scratch_bytes(int n_seg, int max_nq, int max_nt) {
  using namespace sos_hamming_mma;
  const size_t qt = (size_t)((max_nq + 255) / 256) * 2;
  return sos_hamming_mma_scratch_bytes(n_seg, max_nq, max_nt) + (size_t)n_seg * qt * TILE * sizeof(int32_t) + 256;
}

int sos_l2_mma_launch(sos_ctx* ctx, const float* q, const float* t, int dim, const int32_t* q_start, const int32_t* q_len,
                      const int32_t* t_start, const int32_t* t_len, int n_seg, int max_nq, int max_nt, int splits,
                      const uint32_t* items, const int32_t* n_items, unsigned max_items, bool top2, void* exp,
                      ulonglong2* partial, int32_t* flag) {
  using namespace sos_hamming_mma;
  const int qt = ((max_nq + 255) / 256) * 2, tt = (max_nt + TILE - 1) / TILE;
  uint8_t* a_exp = (uint8_t*)exp;
  uint8_t* b_exp = a_exp + (size_t)n_seg * qt * TILE_BYTES;
  int32_t* norms = (int32_t*)(b_exp + sos_align_up((size_t)n_seg * tt * TILE_BYTES, 256));
  if (qt > 0) {
    expand_l2_kernel<<<dim3(qt, n_seg), 256, 0, ctx->stream>>>(q, dim, q_start, q_len, max_nq, qt, 0, a_exp, norms, flag);
    SOS_LAUNCHED_AS(ctx, "l2_expand_kernel");
  }
  if (tt > 0) {
    expand_l2_kernel<<<dim3(tt, n_seg), 256, 0, ctx->stream>>>(t, dim, t_start, t_len, max_nt, tt, 1, b_exp, nullptr, flag);
    SOS_LAUNCHED_AS(ctx, "l2_expand_kernel");
  }
  Args a;
  a.a_norm = norms;
  a.a_exp = a_exp; a.b_exp = b_exp; a.q_len = q_len; a.t_len = t_len;
  a.max_nq = max_nq; a.max_nt = max_nt; a.splits = splits; a.q_tiles_per_seg = qt; a.t_tiles_per_seg = tt;
  a.items = items; a.n_items = n_items; a.partial = partial;
  constexpr int ST = 4;
  static bool attr_set[64][2] = {};
  const int dev = ctx->device & 63, v = top2 ? 1 : 0;
  if (!attr_set[dev][v]) {
    if (top2) SOS_CUDA(cudaFuncSetAttribute(mma_kernel<true, ST, KIND_L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(ST)));
    else SOS_CUDA(cudaFuncSetAttribute(mma_kernel<false, ST, KIND_L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(ST)));
    attr_set[dev][v] = true;
  }
  const bool persistent = max_nt <= 2048;
  const unsigned grid = (persistent && max_items > (unsigned)ctx->sm_count) ? (unsigned)ctx->sm_count : max_items;
  if (top2) mma_kernel<true, ST, KIND_L2><<<grid, THREADS, smem_bytes(ST), ctx->stream>>>(a);
  else mma_kernel<false, ST, KIND_L2><<<grid, THREADS, smem_bytes(ST), ctx->stream>>>(a);
  SOS_LAUNCHED_AS(ctx, "l2_mma_kernel");
  return SOS_OK;
}
