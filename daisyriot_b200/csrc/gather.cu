// gather.cu -- the radiosity / fluorescence gather pass: residual <- M (F residual); B += residual.
//
// Replaces the single-threaded Eigen loops of the reference's Lightning family ("visual studio/Lightning.h"):
//   SpectralLightning::increment_light_fluorescent  :196-226  (K SpMVs + N KxK mat-vecs)
//   RGBLightning::increment_lightpass               :342-349  (K=3, M = diag(rho))
//   BWLightning::increment_lightpass                :419-424  (K=1, M = [1])
//   converge_lightning / check_convergence          :145-151, :255-261, :336-340, :410-415
//
// F is dense FP32, row-major, resident in HBM; one pass streams it exactly once.  Three main kernels, all persistent
// with work items (row block) x (column range):
//   k_gather_tma<K>      K <= 9: a producer warp fills a 6-stage shared-memory ring with tiled TMA loads (F tile +
//                        residual band slices), 8 consumer warps do K FMAs per F element; the default
//   k_gather_mma<K>      K = 16 / 32: tcgen05 3xTF32 tensor-core kernel with the F operand staged in TMEM (gather_mma.cuh)
//   k_gather_partial<K>  the first-generation register-prefetch kernel, kept selectable (DAISY_GATHER=ldg) as a reference point
// k_gather_epilogue sums the column-range partials, applies the per-material KxK matrix, accumulates B, writes the new
// residual into this rank's exchange block -- or, with the fused exchange, into every rank's buffer over NVLink -- and
// totals the per-band residual sums.
#include "daisy_common.cuh"
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define G_WARPS 8
#define G_THREADS (G_WARPS * 32)
#define G_TC 512 // residual tile width (columns)

struct daisy_solver {
    daisy_ctx *ctx = nullptr;
    int K = 0, Kp = 0; // bands requested / bands the kernel is instantiated for
    int nmat = 0;
    int N = 0, n = 0, G = 1, rank = 0, nloc = 0, row0 = 0;
    int64_t bstride = 0;   // floats per exchange block
    int64_t sums_off = 0;  // float offset of the K doubles inside a block
    float *d_res[2] = { nullptr, nullptr }; // exchange buffers (G blocks each)
    int cur = 0;
    float *d_B = nullptr;  // Kp x n   (band-major, local rows)
    float *d_E = nullptr;  // exchange-layout copy of the emission (reset source)
    float *d_M = nullptr;  // nmat x Kp x Kp column-major
    int *d_mat = nullptr;  // local rows
    float *d_partial = nullptr; // nsplit x nloc x Kp
    double *d_cta_sums = nullptr;
    unsigned int *d_done = nullptr;
    unsigned int *d_rb_arrive = nullptr; // fused epilogue (k_gather_tma): column splits finished per row block
    double *d_rb_sums = nullptr;         // per row block band sums
    int nrb_tma = 0;
    unsigned long long *h_abort = nullptr, *d_abort_host = nullptr; // host-mapped abort marker of the fused exchange
    unsigned long long timeout_ns = 30000000000ull;                  // how long a pass waits for a peer's block (DAISY_EXCHANGE_TIMEOUT_MS)
    int R = 8, nsplit = 1, grid = 148, colw = 0;
    int sk_S = 0, sk_L = 0, sk_total = 0, sk_grid = 0, sk_maxp = 1; // stream-K decomposition of the fused TMA kernel (plan())
    bool use_tma = true;
    bool use_mma = false;               // K = 16 / 32: tcgen05 3xTF32 path (gather_mma.cuh)
    bool fused_epi = false;             // k_gather_tma does the epilogue (and the exchange wait) itself: one launch per pass
    float *d_split = nullptr;           // [Rh; Rl]: 2*Kp x ncolsP, rewritten at the start of every pass
    int ncolsP = 0;
    alignas(64) CUtensorMap tmF;        // F rows of this rank: 2-D (ldF x nloc), box 128 x tile rows
    alignas(64) CUtensorMap tmRes[2];   // exchange buffers: 3-D (n x Kp x G), box 128 x Kp x 1
    alignas(64) CUtensorMap tmFmma;     // F rows, box 32 x 128, SWIZZLE_128B
    alignas(64) CUtensorMap tmSplit;    // d_split, box 32 x 2*Kp, SWIZZLE_128B
    // fused exchange (multi-GPU): the epilogue stores this rank's block straight into every rank's next buffer through
    // CUDA-IPC mappings (NVLink), then raises flag[rank] = pass sequence number in every rank's flag array
    bool fused = false;
    bool peers_ipc = false;                            // peer pointers are CUDA-IPC mappings to be closed (one process per GPU)
    float *peer_res[2][16] = { { nullptr } };          // [buffer][rank] base of that rank's exchange buffer
    unsigned long long *peer_flags[16] = { nullptr };  // [rank] base of that rank's flag array
    unsigned long long *d_flags = nullptr;             // own flag array (16 entries) + [16] = wait timeout marker
    unsigned long long seq = 0;                        // passes launched since creation (never reset)
    int numpasses = 0;
    double last_ms = 0.0;
    std::vector<double> sums;   // band sums of the current residual
    bool sums_valid = true;
    std::vector<double> e_sums; // band sums of the emission
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int pin16 = 0;                      // sixteenths of the local F rows kept in L2 across passes (plan())
    bool events_recorded = false;       // e0 / e1 bracket the most recent pass
    bool chained = false;               // daisy_solver_set_chained: passes launched back to back without per-pass events
};

// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// tiled TMA loads through a tensor map (SASS: UTMALDG); out-of-range parts of the box are zero-filled
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// the same with an L2 eviction policy (createpolicy): evict_last keeps a tile in the 126 MB L2 from one pass to the next
__device__ __forceinline__ void tma_load_2d_hint(void *dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float4 ldg_stream(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

struct GatherParams {
    const float *F; int64_t ldF;
    int nloc;               // local rows
    int ncols;              // padded column count G*n
    int n;                  // rows per rank (block length)
    const float *res;       // exchange buffer (read)
    int64_t bstride;
    float *partial;         // nsplit x nloc x K
    int nsplit, colw;       // column range width per split (multiple of G_TC)
    int nrb;                // row blocks
    // stream-K decomposition of k_gather_tma (fused epilogue): the (row block, 256-column step) pairs of the whole matrix form
    // one sequence of sk_total steps, row block major; CTA c streams the contiguous range [c sk_L, (c+1) sk_L) -- every SM
    // gets the same number of steps whatever the matrix shape, and a row block is shared by at most a few CTAs ("pieces")
    int sk;                 // 1: stream-K pieces, 0: (row block, column split) items
    int sk_S, sk_L, sk_total; // steps per row block, steps per CTA, steps in total
    // L2 residency: the F tiles whose step number s has (s * 11 + rb * 5) % 16 < pin16 are loaded evict_last, the others
    // evict_first -- a fixed subset of the matrix, spread evenly over the CTAs, that stays in L2 from pass to pass when the
    // matrix is not much larger than L2 (the reference's own scenes: 164 / 238 MB against 126 MB of L2)
    int pin16;
};

// one unit of work of k_gather_tma: steps [s0, s1) of row block rb; idx / np = this piece's ordinal among / number of the
// pieces that make up the row block (their partial sums are added in idx order)
struct Piece { int rb, s0, s1, idx, np; };

// stage the K bands of columns [j0, j0+w) of the exchange buffer into tile[k][0..w)
template <int K>
__device__ __forceinline__ void issue_tile(const GatherParams &P, float *tile, uint64_t *bar, int j0, int w) {
    mbar_expect_tx(bar, (uint32_t)(K * w * 4));
    int g = j0 / P.n, jl = j0 - g * P.n;
    int w1 = min(w, P.n - jl); // part inside block g; the rest (if any) continues in block g+1
#pragma unroll 1
    for (int k = 0; k < K; k++) {
        bulk_g2s(tile + k * G_TC, P.res + (size_t)g * P.bstride + (size_t)k * P.n + jl, (uint32_t)(w1 * 4), bar);
        if (w1 < w) bulk_g2s(tile + k * G_TC + w1, P.res + (size_t)(g + 1) * P.bstride + (size_t)k * P.n, (uint32_t)((w - w1) * 4), bar);
    }
}

template <int K, int R>
__global__ void __launch_bounds__(G_THREADS, 1) k_gather_partial(GatherParams P) {
    extern __shared__ __align__(128) unsigned char g_smem[];
    float *tiles = reinterpret_cast<float *>(g_smem);                 // 2 x K x G_TC
    uint64_t *bars = reinterpret_cast<uint64_t *>(g_smem + 2 * K * G_TC * sizeof(float));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase[2] = { 0, 0 };
    const int nitems = P.nrb * P.nsplit;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int rb = item / P.nsplit, split = item - rb * P.nsplit;
        const int c_begin = split * P.colw;
        const int c_end = min(P.ncols, c_begin + P.colw);
        const int ntile = (c_end - c_begin + G_TC - 1) / G_TC;
        const int row_base = rb * (G_WARPS * R) + warp * R;
        const float *Frow[R];
#pragma unroll
        for (int r = 0; r < R; r++) Frow[r] = P.F + (size_t)min(row_base + r, P.nloc - 1) * P.ldF + c_begin + lane * 4;
        float acc[R][K];
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int k = 0; k < K; k++) acc[r][k] = 0.0f;
        if (tid == 0) {
            issue_tile<K>(P, tiles, &bars[0], c_begin, min(G_TC, c_end - c_begin));
            if (ntile > 1) issue_tile<K>(P, tiles + K * G_TC, &bars[1], c_begin + G_TC, min(G_TC, c_end - c_begin - G_TC));
        }
        const int nstep = (c_end - c_begin + 127) / 128;
        float4 fcur[R], fnxt[R];
        {
            bool in = (lane * 4) < (c_end - c_begin);
#pragma unroll
            for (int r = 0; r < R; r++) fcur[r] = in ? ldg_stream(Frow[r]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int s = 0; s < nstep; s++) {
            const int t = s >> 2, buf = t & 1;
            if (s + 1 < nstep) {
                bool in = ((s + 1) * 128 + lane * 4) < (c_end - c_begin);
#pragma unroll
                for (int r = 0; r < R; r++) fnxt[r] = in ? ldg_stream(Frow[r] + (size_t)(s + 1) * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if ((s & 3) == 0) { mbar_wait(&bars[buf], phase[buf]); phase[buf] ^= 1; }
            const int col = (s & 3) * 128 + lane * 4; // tile-local column
            if (t * G_TC + col < c_end - c_begin) {
                const float *tile = tiles + buf * K * G_TC + col;
#pragma unroll
                for (int k = 0; k < K; k++) {
                    float4 x = *reinterpret_cast<const float4 *>(tile + k * G_TC);
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        acc[r][k] = fmaf(fcur[r].x, x.x, acc[r][k]);
                        acc[r][k] = fmaf(fcur[r].y, x.y, acc[r][k]);
                        acc[r][k] = fmaf(fcur[r].z, x.z, acc[r][k]);
                        acc[r][k] = fmaf(fcur[r].w, x.w, acc[r][k]);
                    }
                }
            }
            if ((s & 3) == 3 || s == nstep - 1) {
                __syncthreads(); // every warp is done with tile t
                if (tid == 0 && t + 2 < ntile) {
                    int j0 = c_begin + (t + 2) * G_TC;
                    issue_tile<K>(P, tiles + buf * K * G_TC, &bars[buf], j0, min(G_TC, c_end - j0));
                }
            }
#pragma unroll
            for (int r = 0; r < R; r++) fcur[r] = fnxt[r];
        }
        // lane reduction (fixed xor tree => deterministic), lane 0 stores the column-range partial
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int k = 0; k < K; k++) {
                float v = acc[r][k];
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                acc[r][k] = v;
            }
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                int row = row_base + r;
                if (row < P.nloc) {
                    float *dst = P.partial + ((size_t)split * P.nloc + row) * K;
#pragma unroll
                    for (int k = 0; k < K; k++) dst[k] = acc[r][k];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_gather_tma: the same work decomposition with a register-free load pipeline.  Warp 8 is the producer: for every
// 128-column step it fills one shared-memory stage -- the 64x128 F tile as 64 row segments plus the K residual band
// slices -- with 1-D TMA bulk copies that complete on the stage's "full" mbarrier.  Warps 0-7 consume: LDS.128 of their
// 8 F rows and the K band slices, K FMAs per F element, then one arrive on the stage's "empty" mbarrier.  Up to
// NSTAGE x (64+K) x 512 B (~219 KB for K=9) are in flight per SM without holding a single register, and there is no
// CTA-wide barrier in the loop.
// The F tile of the K <= 9 kernel is 64 rows x 256 columns: 1 KB of every row per TMA box, three 73 KB stages.  With 64 x 128
// tiles (512 B per row, six stages -- the first version) the same kernel reached 6.9 TB/s at 131072 columns; 1 KB segments
// give 7.1 TB/s (DRAM page locality).  32 x 256 tiles (five stages) are as fast on one GPU but lose on row-sharded matrices,
// where the residual tile (K x 256 per box, from L2) is a larger share of the traffic (measured on 2 GPUs: 0.371 / 0.357 /
// 0.345 ms per pass for 32x256 / 64x128 / 64x256 at 32768 patches).
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// rows per consumer warp / consumer warps per CTA: the K accumulators of RW rows must fit the register file
template <int K> struct TmaCfg { static constexpr int RW = 8, NCW = 8, COLS = 256; };   // K <= 9 : 64 x 256 tiles, 9 warps
template <> struct TmaCfg<16> { static constexpr int RW = 4, NCW = 16, COLS = 128; };   // 64 x 128 tiles, 17 warps
template <> struct TmaCfg<32> { static constexpr int RW = 4, NCW = 8, COLS = 128; };    // 32 x 128 tiles, 9 warps (FP32-pipe bound anyway)
template <int K>
__host__ __device__ constexpr int tma_rows() { return TmaCfg<K>::RW * TmaCfg<K>::NCW; }
template <int K>
__host__ __device__ constexpr int tma_stage_bytes() { return (tma_rows<K>() + K) * TmaCfg<K>::COLS * 4; }
// shared memory kept back for the fused epilogue: the block's K sums per row (single column split: they never leave the SM)
// and the scratch of the ordered band-sum reductions
template <int K>
__host__ __device__ constexpr int tma_epi_bytes() { return 2 * tma_rows<K>() * K * 4 + 2048 + 64; }
template <int K>
__host__ __device__ constexpr int tma_nstage() {
    return (227 * 1024 - 256 - tma_epi_bytes<K>()) / tma_stage_bytes<K>() > 8 ? 8 : (227 * 1024 - 256 - tma_epi_bytes<K>()) / tma_stage_bytes<K>();
}

// What used to be a second kernel (k_gather_epilogue) and, multi-GPU, a third one in front (k_wait_flags), folded into the
// streaming kernel:
//  * the CTA that completes the LAST column split of a row block (per-block arrival counter) sums the splits in split order,
//    applies the material matrices, accumulates B, stores the block's new residual -- into this rank's exchange block or,
//    fused exchange, into every rank's buffer over NVLink -- and leaves the block's band sums in rb_sums;
//  * the CTA that completes the last row block of the pass totals the band sums in block order and publishes them (and,
//    fused exchange, raises flag[rank] = seq in every rank's flag array);
//  * fused exchange: the producer lane polls flag[g] >= wait_seq only right before the first residual tile of rank g's block,
//    so the stream starts on the blocks that have already arrived (the F tile of that step is already in flight).
// A peer that does not show up within timeout_ns sets the abort markers (device flag word 16 and a host-mapped word): this
// pass then publishes nothing, every later pass returns at once, and the host reports the error at its next call.
struct FusedEpi {
    int enabled;
    unsigned int *rb_arrive;  // per row block: column splits finished (returns to 0 by the end of the pass)
    double *rb_sums;          // nrb x K
    unsigned int *done;       // row blocks finished
    const float *M; const int *mat; float *B;
    float *res_out_block;     // this rank's block of the next exchange buffer (npeers == 0)
    double *block_sums;
    int npeers;
    float *peer_out[16];
    double *peer_sums[16];
    unsigned long long *peer_flag[16];
    unsigned long long seq;
    unsigned long long *flags;       // own flag array; [16] = abort marker
    unsigned long long *host_abort;  // host-mapped copy of the abort marker (may be null)
    int G;
    unsigned long long wait_seq;     // 0: nothing to wait for
    unsigned long long timeout_ns;
};


// the i-th piece of CTA `cta` (cursor: running position, start at 0); false when the CTA has no more work
template <int T_COLS>
__device__ __forceinline__ bool next_piece(const GatherParams &P, int cta, int ncta, int &cursor, Piece &p) {
    if (P.sk) {
        const int g0 = cta * P.sk_L + cursor, g1 = min(P.sk_total, (cta + 1) * P.sk_L);
        if (g0 >= g1) return false;
        p.rb = g0 / P.sk_S;
        p.s0 = g0 - p.rb * P.sk_S;
        p.s1 = min(P.sk_S, p.s0 + (g1 - g0));
        const int first = (p.rb * P.sk_S) / P.sk_L, lastc = ((p.rb + 1) * P.sk_S - 1) / P.sk_L;
        p.idx = cta - first;
        p.np = lastc - first + 1;
        cursor += p.s1 - p.s0;
        return true;
    }
    const int item = cta + cursor * ncta;
    if (item >= P.nrb * P.nsplit) return false;
    p.rb = item / P.nsplit;
    p.idx = item - p.rb * P.nsplit;
    p.np = P.nsplit;
    const int c_begin = p.idx * P.colw, c_end = min(P.ncols, c_begin + P.colw);
    p.s0 = c_begin / T_COLS;
    p.s1 = p.s0 + (c_end - c_begin + T_COLS - 1) / T_COLS;
    cursor++;
    return true;
}

template <int K>
__global__ void __launch_bounds__((TmaCfg<K>::NCW + 3) * 32, 1)
k_gather_tma(GatherParams P, const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ FusedEpi E) {
    constexpr int NST = tma_nstage<K>();
    constexpr int STAGE_F = tma_stage_bytes<K>() / 4; // floats per stage
    constexpr int RW = TmaCfg<K>::RW, NCW = TmaCfg<K>::NCW, T_ROWS = RW * NCW, T_COLS = TmaCfg<K>::COLS;
    extern __shared__ __align__(128) unsigned char g_smem[];
    float *stages = reinterpret_cast<float *>(g_smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(g_smem + (size_t)NST * tma_stage_bytes<K>());
    uint64_t *empty = full + NST;
    float *s_out = reinterpret_cast<float *>(empty + NST + 2);              // 2 x T_ROWS x K sums of a block (single split), double-buffered
    double *s_red = reinterpret_cast<double *>(s_out + 2 * tma_rows<K>() * K); // 2 KB: ordered reductions
    int *s_flag = reinterpret_cast<int *>(empty + NST);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (E.enabled && *reinterpret_cast<volatile unsigned long long *>(E.flags + 16) != 0ull) return; // an earlier pass gave up on a peer
    if (tid == 0) {
        for (int i = 0; i < NST; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // Programmatic dependent launch (chained passes, launch_pass_K): the next pass's grid may be scheduled onto SMs as this one
    // leaves them.  Its producer starts streaming F -- which no pass ever writes -- and only then waits for this grid to
    // complete before touching the residual; everything else in the kernel is downstream of the producer's loads, and the
    // other warps wait as well before they read or write global memory.  Without the launch attribute both are no-ops.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    uint32_t it = 0; // global step counter: stage = it % NST, phase = (it / NST) & 1
    int cursor = 0;
    Piece pc;
    if (warp != NCW) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (warp > NCW) {
        // ------------------------------- epilogue warps (64 threads) -------------------------------
        // They take over each finished item from the consumers (named barriers 2/3: "item done", 4/5: "buffer free"), so
        // the streaming never waits for an epilogue; only the last epilogue of a CTA is exposed.
        if (!E.enabled) return;
        const int et = tid - (NCW + 1) * 32, ewarp = et >> 5;
        auto esync = [&]() { asm volatile("bar.sync 1, 64;" ::: "memory"); };
        unsigned nitem_done = 0;
        for (; next_piece<T_COLS>(P, blockIdx.x, gridDim.x, cursor, pc); nitem_done++) {
            const int rb = pc.rb;
            const int buf = (int)(nitem_done & 1u);
            DZ_ASSERT(rb >= 0 && rb < P.nrb && pc.idx >= 0 && pc.idx < pc.np);
            asm volatile("bar.sync %0, %1;" ::"r"(2 + buf), "n"(NCW * 32 + 64) : "memory"); // the consumers have written this item's sums
            bool last = true;
            if (pc.np > 1) {
                if (et == 0) {
                    // release: the consumers' partial rows (ordered before this thread by the named barrier) become visible
                    // device-wide before the arrival is counted -- barrier + one fence + atomic, the usual semaphore pattern
                    __threadfence();
                    const unsigned old = atomicAdd(&E.rb_arrive[rb], 1u);
                    const int l = (old == (unsigned)pc.np - 1u);
                    if (l) E.rb_arrive[rb] = 0u;
                    *s_flag = l;
                }
                esync();
                last = *s_flag != 0;
                esync();
                if (last) __threadfence();
            }
            if (last) {
                double loc[K];
#pragma unroll
                for (int k = 0; k < K; k++) loc[k] = 0.0;
                const int row = rb * T_ROWS + et;
                if (et < T_ROWS && row < P.nloc) {
                    float b[K], Bv[K];
#pragma unroll
                    for (int k = 0; k < K; k++) Bv[k] = E.B[(size_t)k * P.n + row]; // issued first: independent of everything below
                    DZ_ASSERT(E.mat[row] >= 0 && row < P.n);
                    const float *Mp = E.M + (size_t)E.mat[row] * K * K;
                    if (pc.np == 1) {
#pragma unroll
                        for (int k = 0; k < K; k++) b[k] = s_out[buf * (T_ROWS * K) + et * K + k];
                    } else {
#pragma unroll
                        for (int k = 0; k < K; k++) b[k] = 0.0f;
                        for (int sp = 0; sp < pc.np; sp++) { // fixed order: the result does not depend on which CTA came last
                            const float *src = P.partial + ((size_t)sp * P.nloc + row) * K;
#pragma unroll
                            for (int k = 0; k < K; k++) b[k] += __ldcg(src + k);
                        }
                    }
                    float y[K];
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        // result = reflectionmatrix[i] * patchrowvec  (column-major K x K)            Lightning.h:212
                        float v = 0.0f;
#pragma unroll
                        for (int j = 0; j < K; j++) v = fmaf(__ldg(Mp + j * K + k), b[j], v);
                        y[k] = v;
                        loc[k] = (double)v;
                    }
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        if (E.npeers > 0) {
                            for (int g = 0; g < E.npeers; g++) E.peer_out[g][(size_t)k * P.n + row] = y[k]; // NVLink stores into every rank's buffer
                        } else {
                            E.res_out_block[(size_t)k * P.n + row] = y[k];                   // residualvector[j][i] = result[j]
                        }
                        E.B[(size_t)k * P.n + row] = Bv[k] + y[k];                           // lightningvalues += residual   :221-223
                    }
                    if (E.npeers > 0) __threadfence_system(); else __threadfence();
                }
                // band sums of the block: lanes -> warps -> block, always in index order
#pragma unroll
                for (int k = 0; k < K; k++) {
                    double v = loc[k];
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0) s_red[ewarp * K + k] = v;
                }
                esync();
                if (et < K) {
                    E.rb_sums[(size_t)rb * K + et] = s_red[et] + s_red[K + et];
                    __threadfence();
                }
                esync();
                if (et == 0) *s_flag = (atomicAdd(E.done, 1u) == (unsigned)P.nrb - 1u);
                esync();
                const bool final_block = *s_flag != 0;
                esync();
                if (final_block) {
                    // the pass's last row block: total the band sums in block order (a fixed two-level order) and publish
                    __threadfence();
                    constexpr int NCH = (256 / K) < 32 ? (256 / K) : 32; // chunks of consecutive row blocks per band (NCH x K doubles fit s_red)
                    const int per = (P.nrb + NCH - 1) / NCH;
                    for (int t = et; t < NCH * K; t += 64) {
                        const int k = t % K, c = t / K;
                        double v = 0.0;
                        const volatile double *rs = E.rb_sums;
                        for (int b2 = c * per; b2 < min(P.nrb, (c + 1) * per); b2++) v += rs[(size_t)b2 * K + k];
                        s_red[c * K + k] = v;
                    }
                    esync();
                    const bool aborted = *reinterpret_cast<volatile unsigned long long *>(E.flags + 16) != 0ull;
                    if (et < K) {
                        double v = 0.0;
                        for (int c = 0; c < NCH; c++) v += s_red[c * K + et];
                        if (E.npeers > 0) { for (int g = 0; g < E.npeers; g++) E.peer_sums[g][et] = v; }
                        else E.block_sums[et] = v;
                    }
                    if (et == 0) *E.done = 0u;
                    if (E.npeers > 0) {
                        __threadfence_system();
                        esync();
                        if (et < E.npeers && !aborted) {
                            // release: every store of this pass (fenced by its CTA before it arrived on `done`) precedes the flag
                            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(E.peer_flag[et]), "l"(E.seq) : "memory");
                        }
                    }
                    esync();
                }
            }
            asm volatile("bar.arrive %0, %1;" ::"r"(4 + buf), "n"(NCW * 32 + 64) : "memory"); // this buffer may be refilled
        }
    } else if (warp == NCW) {
        // ------------------------------- producer -------------------------------
        // one elected lane issues two tiled TMA loads per stage: the T_ROWS x 128 block of F and the K x 128 block
        // of the residual bands (tiles never straddle an exchange block: n is a multiple of 256 when G > 1, and a
        // single block's tail beyond n is zero-filled by the TMA unit)
        if (lane == 0) {
            unsigned ready = (E.enabled && E.wait_seq != 0ull) ? 0u : 0xffffffffu; // ranks whose block of the previous pass is known to be here
            uint64_t pol_last, pol_first;
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
            auto pol = [&](int rb, int s) -> uint64_t { return (((s * 11 + rb * 5) & 15) < P.pin16) ? pol_last : pol_first; };
            // the F tiles of the first NST steps go out before the previous pass is known to be complete
            uint32_t pre = 0;
            {
                int cur2 = 0;
                Piece p2;
                while (pre < (uint32_t)NST && next_piece<T_COLS>(P, blockIdx.x, gridDim.x, cur2, p2))
                    for (int s = p2.s0; s < p2.s1 && pre < (uint32_t)NST; s++, pre++) {
                        mbar_expect_tx(&full[pre], (uint32_t)tma_stage_bytes<K>());
                        tma_load_2d_hint(stages + (size_t)pre * STAGE_F, &tmF, s * T_COLS, p2.rb * T_ROWS, &full[pre], pol(p2.rb, s));
                    }
            }
            asm volatile("griddepcontrol.wait;" ::: "memory");
            while (next_piece<T_COLS>(P, blockIdx.x, gridDim.x, cursor, pc)) {
                const int row0 = pc.rb * T_ROWS;
                for (int s = pc.s0; s < pc.s1; s++, it++) {
                    const int st = it % NST;
                    const int col0 = s * T_COLS;
                    float *sf = stages + (size_t)st * STAGE_F;
                    if (it >= pre) {
                        mbar_wait(&empty[st], ((it / NST) & 1) ^ 1);
                        mbar_expect_tx(&full[st], (uint32_t)tma_stage_bytes<K>());
                        tma_load_2d_hint(sf, &tmF, col0, row0, &full[st], pol(pc.rb, s));
                    }
                    const int g = col0 / P.n, jl = col0 - g * P.n;
                    DZ_ASSERT(g >= 0 && g < 32 && col0 < P.ncols && row0 < P.nloc + T_ROWS);
                    if (!((ready >> g) & 1u)) {
                        unsigned long long t0, t1, v;
                        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
                        while (true) {
                            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(E.flags + g) : "memory");
                            if (v >= E.wait_seq) break;
                            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
                            if (t1 - t0 > E.timeout_ns) {
                                E.flags[16] = E.seq;
                                if (E.host_abort) *reinterpret_cast<volatile unsigned long long *>(E.host_abort) = E.seq;
                                __threadfence_system();
                                break;
                            }
                            __nanosleep(100);
                        }
                        ready |= 1u << g;
                        asm volatile("fence.proxy.async;" ::: "memory"); // the peer's stores, seen by the acquire above, before the TMA reads below
                    }
                    tma_load_3d(sf + T_ROWS * T_COLS, &tmRes, jl, 0, g, &full[st]);
                }
            }
        }
    } else {
        // ------------------------------- consumers ------------------------------
        unsigned nitem_done = 0;
        while (next_piece<T_COLS>(P, blockIdx.x, gridDim.x, cursor, pc)) {
            const int split = pc.idx;
            const int row_base = pc.rb * T_ROWS + warp * RW;
            float acc[RW][K];
#pragma unroll
            for (int r = 0; r < RW; r++)
#pragma unroll
                for (int k = 0; k < K; k++) acc[r][k] = 0.0f;
            for (int s = pc.s0; s < pc.s1; s++, it++) {
                const int st = it % NST;
                mbar_wait(&full[st], (it / NST) & 1);
#pragma unroll
                for (int hcol = 0; hcol < T_COLS; hcol += 128) { // a lane covers 4 columns of every 128
                    const float *sf = stages + (size_t)st * STAGE_F + lane * 4 + hcol;
                    float4 f[RW];
#pragma unroll
                    for (int r = 0; r < RW; r++) f[r] = *reinterpret_cast<const float4 *>(sf + (warp * RW + r) * T_COLS);
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        float4 x = *reinterpret_cast<const float4 *>(sf + (T_ROWS + k) * T_COLS);
#pragma unroll
                        for (int r = 0; r < RW; r++) {
                            acc[r][k] = fmaf(f[r].x, x.x, acc[r][k]);
                            acc[r][k] = fmaf(f[r].y, x.y, acc[r][k]);
                            acc[r][k] = fmaf(f[r].z, x.z, acc[r][k]);
                            acc[r][k] = fmaf(f[r].w, x.w, acc[r][k]);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[st]);
            }
#pragma unroll
            for (int r = 0; r < RW; r++)
#pragma unroll
                for (int k = 0; k < K; k++) {
                    float v = acc[r][k];
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    acc[r][k] = v;
                }
            // hand the block's sums to the epilogue warps (double-buffered: the consumers go straight on to the next item)
            const int buf = (int)(nitem_done & 1u);
            if (E.enabled && nitem_done >= 2u) asm volatile("bar.sync %0, %1;" ::"r"(4 + buf), "n"(NCW * 32 + 64) : "memory"); // epilogue of item - 2 has let go of this buffer
            if (lane == 0) {
#pragma unroll
                for (int r = 0; r < RW; r++) {
                    int row = row_base + r;
                    if (E.enabled && pc.np == 1) {
                        float *dst = s_out + buf * (T_ROWS * K) + (warp * RW + r) * K;
#pragma unroll
                        for (int k = 0; k < K; k++) dst[k] = acc[r][k];
                    } else if (row < P.nloc) {
                        float *dst = P.partial + ((size_t)split * P.nloc + row) * K;
#pragma unroll
                        for (int k = 0; k < K; k++) dst[k] = acc[r][k];
                    }
                }
            }
            if (E.enabled) asm volatile("bar.arrive %0, %1;" ::"r"(2 + buf), "n"(NCW * 32 + 64) : "memory");
            nitem_done++;
        }
    }
}

#include "gather_mma.cuh"

struct EpiParams {
    const float *partial; int nsplit, nloc, n;
    const float *M; const int *mat;
    float *res_out_block; // this rank's block of the next exchange buffer: K x n floats
    float *B;             // K x n
    double *cta_sums;     // gridDim x K
    double *block_sums;   // K doubles in the tail of this rank's block
    unsigned int *done;
    // fused exchange: npeers > 0 => the block (and its band sums) go to peer_out[g] / peer_sums[g] of every rank g
    // (own rank included) and, once every CTA's stores are fenced, flag[rank] of every rank is set to seq
    int npeers;
    float *peer_out[16];
    double *peer_sums[16];
    unsigned long long *peer_flag[16];
    unsigned long long seq;
    const unsigned long long *abort_marker; // flag word 16 of this rank (null: single GPU)
};

template <int K>
__global__ void __launch_bounds__(256) k_gather_epilogue(EpiParams P) {
    __shared__ double s_sum[8][K];
    __shared__ bool s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double loc[K];
#pragma unroll
    for (int k = 0; k < K; k++) loc[k] = 0.0;
    for (int row = blockIdx.x * blockDim.x + tid; row < P.nloc; row += gridDim.x * blockDim.x) {
        float b[K];
#pragma unroll
        for (int k = 0; k < K; k++) b[k] = 0.0f;
        for (int s = 0; s < P.nsplit; s++) {
            const float *src = P.partial + ((size_t)s * P.nloc + row) * K;
#pragma unroll
            for (int k = 0; k < K; k++) b[k] += src[k];
        }
        const float *Mp = P.M + (size_t)P.mat[row] * K * K;
#pragma unroll
        for (int k = 0; k < K; k++) {
            // result = reflectionmatrix[i] * patchrowvec  (column-major K x K)            Lightning.h:212
            float y = 0.0f;
#pragma unroll
            for (int j = 0; j < K; j++) y = fmaf(Mp[j * K + k], b[j], y);
            if (P.npeers > 0) {
                for (int g = 0; g < P.npeers; g++) P.peer_out[g][(size_t)k * P.n + row] = y; // NVLink stores into every rank's buffer
            } else {
                P.res_out_block[(size_t)k * P.n + row] = y;                      // residualvector[j][i] = result[j]
            }
            P.B[(size_t)k * P.n + row] = P.B[(size_t)k * P.n + row] + y;         // lightningvalues += residual   :221-223
            loc[k] += (double)y;
        }
    }
    // deterministic totals: lanes -> warps -> CTA -> (last CTA) grid, always in index order
#pragma unroll
    for (int k = 0; k < K; k++) {
        double v = loc[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_sum[warp][k] = v;
    }
    __syncthreads();
    if (tid < K) {
        double v = 0.0;
        for (int w = 0; w < 8; w++) v += s_sum[w][tid];
        P.cta_sums[(size_t)blockIdx.x * K + tid] = v;
    }
    if (P.npeers > 0) __threadfence_system(); else __threadfence(); // peer stores must be visible before the flag below
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(P.done, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last) {
        __threadfence();
        if (tid < K) {
            double v = 0.0;
            volatile const double *cs = P.cta_sums;
            for (unsigned b = 0; b < gridDim.x; b++) v += cs[(size_t)b * K + tid];
            if (P.npeers > 0) { for (int g = 0; g < P.npeers; g++) P.peer_sums[g][tid] = v; }
            else P.block_sums[tid] = v;
        }
        if (tid == 0) *P.done = 0;
        if (P.npeers > 0) {
            __threadfence_system();
            __syncthreads();
            const bool aborted = P.abort_marker && *reinterpret_cast<const volatile unsigned long long *>(P.abort_marker) != 0ull;
            if (tid < P.npeers && !aborted) {
                // release: every store of this kernel (fenced by each CTA before it arrived on `done`) precedes the flag
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(P.peer_flag[tid]), "l"(P.seq) : "memory");
            }
        }
    }
}

// waits until every rank's block of pass `seq` has arrived in this rank's buffer (flag[g] >= seq for all g).  A peer that does
// not show up within timeout_ns turns into an error instead of a hung GPU: the abort markers are set (flag word 16 and the
// host-mapped word), the kernels of later passes return at once and nothing more is published to the peers.
__global__ void k_wait_flags(unsigned long long *flags, int G, unsigned long long seq, unsigned long long timeout_ns, unsigned long long *host_abort) {
    const int g = threadIdx.x;
    if (g >= G) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    while (true) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + g) : "memory");
        if (v >= seq) break;
        if (*reinterpret_cast<volatile unsigned long long *>(flags + 16) != 0ull) break;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) {
            flags[16] = seq;
            if (host_abort) *reinterpret_cast<volatile unsigned long long *>(host_abort) = seq;
            __threadfence_system();
            break;
        }
        __nanosleep(200);
    }
}

// ---------------------------------------------------------------------------------------------------------
template <int K, int R>
static int launch_partial(daisy_solver *s, const GatherParams &P) {
    size_t smem = 2 * K * G_TC * sizeof(float) + 64;
    static bool attr_done[64] = { false }; // per device: a function attribute belongs to the device's context
    if (!attr_done[s->ctx->device & 63]) {
        DZ_CUDA(cudaFuncSetAttribute(k_gather_partial<K, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done[s->ctx->device & 63] = true;
    }
    k_gather_partial<K, R><<<s->grid, G_THREADS, smem, s->ctx->stream>>>(P);
    DZ_CUDA(cudaGetLastError());
    return DAISY_OK;
}

// wait_seq != 0 (fused exchange): the blocks of pass wait_seq must have arrived before the residual is read.  The fused TMA
// kernel waits by itself, block by block; the other kernels get a k_wait_flags launch in front.
template <int K>
static int launch_pass_K(daisy_solver *s, unsigned long long wait_seq) {
    daisy_ctx *c = s->ctx;
    if (wait_seq != 0ull && !s->fused_epi) {
        k_wait_flags<<<1, 32, 0, c->stream>>>(s->d_flags, s->G, wait_seq, s->timeout_ns, s->d_abort_host);
        DZ_CUDA(cudaGetLastError());
    }
    GatherParams P;
    P.F = c->d_F; P.ldF = c->ldF; P.nloc = s->nloc; P.ncols = s->G * s->n; P.n = s->n;
    P.res = s->d_res[s->cur]; P.bstride = s->bstride; P.partial = s->d_partial;
    P.nsplit = s->nsplit; P.colw = s->colw; P.nrb = (s->nloc + G_WARPS * s->R - 1) / (G_WARPS * s->R);
    P.sk = 0; P.sk_S = P.sk_L = P.sk_total = 0; P.pin16 = 0;
    int rc;
    if (s->use_mma && (K == 16 || K == 32)) {
        if constexpr (K == 16 || K == 32) {
            constexpr size_t smem = mm_smem_bytes<K>();
            static bool attr_done[64] = { false }; // per device
            if (!attr_done[c->device & 63]) {
                DZ_CUDA(cudaFuncSetAttribute(k_gather_mma<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                attr_done[c->device & 63] = true;
            }
            k_split_residual<K><<<(s->ncolsP + 255) / 256, 256, 0, c->stream>>>(s->d_res[s->cur], s->bstride, s->n, s->G * s->n, s->ncolsP, s->d_split);
            DZ_CUDA(cudaGetLastError());
            P.nrb = (s->nloc + MM_ROWS - 1) / MM_ROWS;
            k_gather_mma<K><<<s->grid, MM_THREADS, smem, c->stream>>>(P, s->tmFmma, s->tmSplit);
            DZ_CUDA(cudaGetLastError());
        }
        rc = DAISY_OK;
    } else if (s->use_tma) {
        constexpr int NST = tma_nstage<K>();
        size_t smem = (size_t)NST * tma_stage_bytes<K>() + 2 * NST * sizeof(uint64_t) + tma_epi_bytes<K>();
        static bool attr_done[64] = { false }; // per device
        if (!attr_done[c->device & 63]) {
            DZ_CUDA(cudaFuncSetAttribute(k_gather_tma<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_done[c->device & 63] = true;
        }
        P.nrb = (s->nloc + tma_rows<K>() - 1) / tma_rows<K>();
        P.sk = s->fused_epi ? 1 : 0; P.sk_S = s->sk_S; P.sk_L = s->sk_L; P.sk_total = s->sk_total;
        P.pin16 = s->pin16;
        const int grid = s->fused_epi ? s->sk_grid : s->grid;
        FusedEpi F;
        memset(&F, 0, sizeof(F));
        F.enabled = s->fused_epi ? 1 : 0;
        if (s->fused_epi) {
            float *next = s->d_res[s->cur ^ 1] + (size_t)s->rank * s->bstride;
            F.rb_arrive = s->d_rb_arrive; F.rb_sums = s->d_rb_sums; F.done = s->d_done;
            F.M = s->d_M; F.mat = s->d_mat; F.B = s->d_B;
            F.res_out_block = next; F.block_sums = reinterpret_cast<double *>(next + s->sums_off);
            F.seq = s->seq; F.flags = s->d_flags; F.host_abort = s->d_abort_host; F.G = s->G;
            F.wait_seq = wait_seq; F.timeout_ns = s->timeout_ns;
            if (s->fused) {
                F.npeers = s->G;
                for (int g = 0; g < s->G; g++) {
                    float *blk = s->peer_res[s->cur ^ 1][g] + (size_t)s->rank * s->bstride;
                    F.peer_out[g] = blk;
                    F.peer_sums[g] = reinterpret_cast<double *>(blk + s->sums_off);
                    F.peer_flag[g] = s->peer_flags[g] + s->rank;
                }
            }
        }
        {
            // chained single-GPU passes: programmatic dependent launch lets this grid move onto SMs as the previous pass leaves
            // them and stream its first F tiles meanwhile (the kernel waits for the previous grid before it touches anything
            // a pass writes).  Multi-GPU passes keep plain stream order (their abort marker is read at kernel start).
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((TmaCfg<K>::NCW + 3) * 32); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at;
            cfg.numAttrs = (s->fused_epi && s->chained && s->G == 1) ? 1 : 0;
            DZ_CUDA(cudaLaunchKernelEx(&cfg, k_gather_tma<K>, P, s->tmF, s->tmRes[s->cur], F));
        }
        DZ_CUDA(cudaGetLastError());
        if (s->fused_epi) return DAISY_OK; // the whole pass was that one launch
        rc = DAISY_OK;
    } else if constexpr (K <= 9) {
        rc = (s->R == 8) ? launch_partial<K, 8>(s, P) : launch_partial<K, 4>(s, P);
    } else {
        rc = (s->R == 4) ? launch_partial<K, 4>(s, P) : launch_partial<K, 2>(s, P);
    }
    if (rc) return rc;
    EpiParams E;
    float *next = s->d_res[s->cur ^ 1] + (size_t)s->rank * s->bstride;
    E.partial = s->d_partial; E.nsplit = s->nsplit; E.nloc = s->nloc; E.n = s->n;
    E.M = s->d_M; E.mat = s->d_mat; E.res_out_block = next; E.B = s->d_B;
    E.cta_sums = s->d_cta_sums; E.block_sums = reinterpret_cast<double *>(next + s->sums_off); E.done = s->d_done;
    E.npeers = 0; E.seq = s->seq; E.abort_marker = s->fused ? s->d_flags + 16 : nullptr;
    if (s->fused) {
        E.npeers = s->G;
        for (int g = 0; g < s->G; g++) {
            float *blk = s->peer_res[s->cur ^ 1][g] + (size_t)s->rank * s->bstride;
            E.peer_out[g] = blk;
            E.peer_sums[g] = reinterpret_cast<double *>(blk + s->sums_off);
            E.peer_flag[g] = s->peer_flags[g] + s->rank;
        }
    }
    int egrid = (s->nloc + 255) / 256;
    if (egrid > 148) egrid = 148;
    if (egrid < 1) egrid = 1;
    k_gather_epilogue<K><<<egrid, 256, 0, c->stream>>>(E);
    DZ_CUDA(cudaGetLastError());
    return DAISY_OK;
}

static int launch_pass(daisy_solver *s, unsigned long long wait_seq = 0ull) {
    switch (s->Kp) {
    case 1: return launch_pass_K<1>(s, wait_seq);
    case 3: return launch_pass_K<3>(s, wait_seq);
    case 9: return launch_pass_K<9>(s, wait_seq);
    case 16: return launch_pass_K<16>(s, wait_seq);
    case 32: return launch_pass_K<32>(s, wait_seq);
    }
    daisy_set_error("unsupported padded band count %d", s->Kp);
    return DAISY_E_INVALID;
}

// ---- tensor maps (driver entry point fetched through the runtime; no libcuda link dependency) ------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

static int make_maps(daisy_solver *s) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) { daisy_set_error("cuTensorMapEncodeTiled is not available from this driver"); return DAISY_E_CUDA; }
    daisy_ctx *c = s->ctx;
    const int trows = (s->Kp == 32) ? 32 : 64, tcols = (s->Kp <= 9) ? 256 : 128; // = tma_rows<Kp>() x TmaCfg<Kp>::COLS
    {
        cuuint64_t dims[2] = { (cuuint64_t)c->ldF, (cuuint64_t)(s->nloc > 0 ? s->nloc : 1) };
        cuuint64_t strides[1] = { (cuuint64_t)c->ldF * 4 };
        cuuint32_t box[2] = { (cuuint32_t)tcols, (cuuint32_t)trows };
        cuuint32_t es[2] = { 1, 1 };
        CUresult r = enc(&s->tmF, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, c->d_F, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { daisy_set_error("cuTensorMapEncodeTiled(F) failed: %d", (int)r); return DAISY_E_CUDA; }
    }
    for (int b = 0; b < 2; b++) {
        cuuint64_t dims[3] = { (cuuint64_t)s->n, (cuuint64_t)s->Kp, (cuuint64_t)s->G };
        cuuint64_t strides[2] = { (cuuint64_t)s->n * 4, (cuuint64_t)s->bstride * 4 };
        cuuint32_t box[3] = { (cuuint32_t)tcols, (cuuint32_t)s->Kp, 1 };
        cuuint32_t es[3] = { 1, 1, 1 };
        CUresult r = enc(&s->tmRes[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, s->d_res[b], dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { daisy_set_error("cuTensorMapEncodeTiled(residual) failed: %d", (int)r); return DAISY_E_CUDA; }
    }
    if (s->use_mma) {
        {
            cuuint64_t dims[2] = { (cuuint64_t)c->ldF, (cuuint64_t)(s->nloc > 0 ? s->nloc : 1) };
            cuuint64_t strides[1] = { (cuuint64_t)c->ldF * 4 };
            cuuint32_t box[2] = { MM_SUB, MM_ROWS };
            cuuint32_t es[2] = { 1, 1 };
            CUresult r = enc(&s->tmFmma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, c->d_F, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { daisy_set_error("cuTensorMapEncodeTiled(F, swizzled) failed: %d", (int)r); return DAISY_E_CUDA; }
        }
        {
            cuuint64_t dims[2] = { (cuuint64_t)s->ncolsP, (cuuint64_t)(2 * s->Kp) };
            cuuint64_t strides[1] = { (cuuint64_t)s->ncolsP * 4 };
            cuuint32_t box[2] = { MM_SUB, (cuuint32_t)(2 * s->Kp) };
            cuuint32_t es[2] = { 1, 1 };
            CUresult r = enc(&s->tmSplit, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, s->d_split, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { daisy_set_error("cuTensorMapEncodeTiled(split residual) failed: %d", (int)r); return DAISY_E_CUDA; }
        }
    }
    return DAISY_OK;
}

// choose rows-per-warp and the column split so that (row blocks x splits) fills the persistent grid evenly
static void plan(daisy_solver *s) {
    int sms = s->ctx->num_sms;
    s->grid = sms;
    const char *env = getenv("DAISY_GATHER");
    s->use_tma = !(env && strcmp(env, "ldg") == 0);
    // wide band counts run on the tensor cores unless DAISY_GATHER=fp32 (or ldg) asks for the CUDA-core kernels
    s->use_mma = s->use_tma && s->Kp >= 16 && !(env && strcmp(env, "fp32") == 0);
    {
        const char *e2 = getenv("DAISY_GATHER_EPI"); // "split": keep the separate epilogue kernel (A/B switch)
        s->fused_epi = s->use_tma && !s->use_mma && !(e2 && strcmp(e2, "split") == 0);
    }
    int Rs[2];
    if (s->use_mma) { Rs[0] = Rs[1] = MM_ROWS / G_WARPS; } // 128-row tiles
    else if (s->use_tma) { Rs[0] = Rs[1] = (s->Kp == 32 ? 4 : 8); } // rows per block / G_WARPS: 64-row tiles, 32-row for K=32
    else if (s->Kp <= 9) { Rs[0] = 8; Rs[1] = 4; } else { Rs[0] = 2; Rs[1] = 4; }
    int ncols = s->G * s->n;
    // keep at least 2048 columns per item -- 1024 with the fused epilogue, where a split costs one partial row per block and no
    // extra launch: small matrices (the reference's own scenes: ~120 row blocks) then fill the 148 SMs evenly
    const int mincols = s->fused_epi ? 2 * G_TC : 4 * G_TC;
    int maxsplit = (ncols + mincols - 1) / mincols;
    if (maxsplit < 1) maxsplit = 1;
    if (maxsplit > 32) maxsplit = 32;
    double best = -1.0;
    for (int ri = 0; ri < 2; ri++) {
        int R = Rs[ri];
        int nrb = (s->nloc + G_WARPS * R - 1) / (G_WARPS * R);
        for (int sp = 1; sp <= maxsplit; sp++) {
            long items = (long)nrb * sp;
            long rounds = (items + sms - 1) / sms;
            double eff = (double)items / (double)(rounds * sms);
            eff -= 0.004 * (sp - 1) + (ri == 1 ? 0.02 : 0.0); // prefer fewer splits and the first R on ties
            if (eff > best) { best = eff; s->R = R; s->nsplit = sp; }
        }
    }
    int colw = (ncols + s->nsplit - 1) / s->nsplit;
    s->colw = ((colw + G_TC - 1) / G_TC) * G_TC; // multiple of 512 columns (and of the 128-column TMA step)
    s->nsplit = (ncols + s->colw - 1) / s->colw;
    if (s->fused_epi) {
        // stream-K: the (row block, column step) sequence cut into one contiguous, equally long range per SM
        const int trows = (s->Kp == 32) ? 32 : 64, tcols = (s->Kp <= 9) ? 256 : 128; // = tma_rows<Kp>() x TmaCfg<Kp>::COLS
        const int nrb = (s->nloc + trows - 1) / trows;
        s->sk_S = (ncols + tcols - 1) / tcols;
        s->sk_total = nrb * s->sk_S;
        s->sk_grid = s->sk_total < sms ? (s->sk_total > 0 ? s->sk_total : 1) : sms;
        s->sk_L = (s->sk_total + s->sk_grid - 1) / s->sk_grid;
        if (s->sk_L < 1) s->sk_L = 1;
        s->sk_maxp = s->sk_S / s->sk_L + 2;
        // L2 residency (GatherParams::pin16): keep ~45 % of the L2 filled with F tiles that are read again next pass; the rest
        // of the matrix streams through evict_first (which by itself is worth 9 % on the 164 MB matrix: the stream no longer
        // pushes the residual vectors out).  Measured on the reference's scenes (164 / 238 MB): best at 61 / 45 MB pinned, worse
        // from ~70 MB on (the two L2 halves mirror lines read from both dies).  DAISY_GATHER_PIN16 = 0..16 overrides.
        {
            int l2 = 0;
            cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, s->ctx->device);
            const double fbytes = (double)s->nloc * (double)ncols * 4.0;
            int p16 = fbytes > 0.0 ? (int)floor(16.0 * 0.45 * (double)l2 / fbytes) : 0;
            if (p16 > 16) p16 = 16;
            if (p16 < 0) p16 = 0;
            const char *e3 = getenv("DAISY_GATHER_PIN16");
            if (e3) { p16 = atoi(e3); if (p16 < 0) p16 = 0; if (p16 > 16) p16 = 16; }
            s->pin16 = p16;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
static int padded_K(int K) { return K <= 1 ? 1 : K <= 3 ? 3 : K <= 9 ? 9 : K <= 16 ? 16 : 32; }

static void free_solver(daisy_solver *s) {
    if (!s) return;
    if (s->fused && s->peers_ipc)
        for (int g = 0; g < s->G; g++) {
            if (g == s->rank) continue;
            if (s->peer_res[0][g]) cudaIpcCloseMemHandle(s->peer_res[0][g]);
            if (s->peer_res[1][g]) cudaIpcCloseMemHandle(s->peer_res[1][g]);
            if (s->peer_flags[g]) cudaIpcCloseMemHandle(s->peer_flags[g]);
        }
    cudaFree(s->d_flags);
    cudaFree(s->d_res[0]); cudaFree(s->d_res[1]); cudaFree(s->d_B); cudaFree(s->d_E); cudaFree(s->d_M); cudaFree(s->d_mat);
    cudaFree(s->d_partial); cudaFree(s->d_cta_sums); cudaFree(s->d_done); cudaFree(s->d_split);
    cudaFree(s->d_rb_arrive); cudaFree(s->d_rb_sums);
    if (s->h_abort) cudaFreeHost(s->h_abort);
    if (s->e0) cudaEventDestroy(s->e0);
    if (s->e1) cudaEventDestroy(s->e1);
    delete s;
}

// host (K x N band-major) -> exchange layout (G blocks of [Kp x n] + tail)
static void to_exchange(const daisy_solver *s, const float *src, std::vector<float> &dst) {
    dst.assign((size_t)s->G * s->bstride, 0.0f);
    for (int k = 0; k < s->K; k++)
        for (int p = 0; p < s->N; p++) {
            int g = p / s->n, jl = p - g * s->n;
            dst[(size_t)g * s->bstride + (size_t)k * s->n + jl] = src[(size_t)k * s->N + p];
        }
}

// FP64 sum of n floats with eight independent partial sums (fixed order, so deterministic): a single dependent chain costs
// 1 ns per element, which was 1.2 of the 1.5 ms a host-buffer step spent outside the pass at 131072 patches x 9 bands
static double host_sum(const float *x, size_t n) {
    double a[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    size_t i = 0;
    for (; i + 8 <= n; i += 8)
        for (int j = 0; j < 8; j++) a[j] += (double)x[i + j];
    for (int j = 0; i < n; i++, j++) a[j] += (double)x[i];
    return ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
}

static int compute_sums_from_host(daisy_solver *s, const float *res /* K x N */) {
    // band sums of a residual supplied by the host (reset / write): same quantity the kernels total on device
    s->sums.assign(s->K, 0.0);
    for (int k = 0; k < s->K; k++) s->sums[k] = host_sum(res + (size_t)k * s->N, (size_t)s->N);
    return DAISY_OK;
}

extern "C" int daisy_solver_create(daisy_ctx *ctx, int K, const float *E, const float *M, int nmat, const int32_t *mat_idx,
                                   daisy_solver **out) {
    DZ_REQUIRE(ctx && out && E && M && mat_idx, DAISY_E_INVALID, "daisy_solver_create: null argument");
    DZ_REQUIRE(K >= 1 && K <= DAISY_MAX_BANDS, DAISY_E_INVALID, "daisy_solver_create: K must be in [1,32]");
    DZ_REQUIRE(nmat >= 1, DAISY_E_INVALID, "daisy_solver_create: need at least one material");
    DZ_REQUIRE(ctx->have_F, DAISY_E_STATE, "daisy_solver_create: build or load the form factors first");
    DZ_REQUIRE(ctx->N > 0, DAISY_E_INVALID, "daisy_solver_create: empty mesh");
    for (int p = 0; p < ctx->N; p++)
        DZ_REQUIRE(mat_idx[p] >= 0 && mat_idx[p] < nmat, DAISY_E_INVALID, "daisy_solver_create: material index out of range");
    DZ_CUDA(cudaSetDevice(ctx->device));
    daisy_solver *s = new daisy_solver();
    s->ctx = ctx; s->K = K; s->Kp = padded_K(K); s->nmat = nmat;
    s->N = ctx->N; s->n = ctx->rows_per_rank; s->G = ctx->nranks; s->rank = ctx->rank;
    s->row0 = ctx->row0; s->nloc = ctx->row1 - ctx->row0;
    int64_t body = (int64_t)s->Kp * s->n;
    s->sums_off = body;
    s->bstride = ((body + 2 * s->Kp + 3) / 4) * 4; // K doubles = 2K floats, block padded to 16 B
    plan(s);
    size_t exb = sizeof(float) * (size_t)s->G * s->bstride;
    int rc = DAISY_OK;
#define SC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { daisy_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); free_solver(s); return DAISY_E_CUDA; } } while (0)
    SC(cudaMalloc(&s->d_res[0], exb));
    SC(cudaMalloc(&s->d_res[1], exb));
    SC(cudaMalloc(&s->d_E, exb));
    SC(cudaMalloc(&s->d_B, sizeof(float) * (size_t)s->Kp * s->n));
    SC(cudaMalloc(&s->d_M, sizeof(float) * (size_t)nmat * s->Kp * s->Kp));
    SC(cudaMalloc(&s->d_mat, sizeof(int) * (size_t)(s->nloc > 0 ? s->nloc : 1)));
    SC(cudaMalloc(&s->d_partial, sizeof(float) * (size_t)(s->nsplit > s->sk_maxp ? s->nsplit : s->sk_maxp) * (s->nloc > 0 ? s->nloc : 1) * s->Kp));
    SC(cudaMalloc(&s->d_cta_sums, sizeof(double) * 148 * s->Kp));
    SC(cudaMalloc(&s->d_done, sizeof(unsigned int)));
    SC(cudaMemset(s->d_done, 0, sizeof(unsigned int)));
    {
        const int trows = (s->Kp == 32) ? 32 : 64; // = tma_rows<Kp>()
        s->nrb_tma = (s->nloc + trows - 1) / trows;
        if (s->nrb_tma < 1) s->nrb_tma = 1;
        SC(cudaMalloc(&s->d_rb_arrive, sizeof(unsigned int) * (size_t)s->nrb_tma));
        SC(cudaMemset(s->d_rb_arrive, 0, sizeof(unsigned int) * (size_t)s->nrb_tma));
        SC(cudaMalloc(&s->d_rb_sums, sizeof(double) * (size_t)s->nrb_tma * s->Kp));
        SC(cudaHostAlloc(reinterpret_cast<void **>(&s->h_abort), sizeof(unsigned long long), cudaHostAllocMapped));
        *s->h_abort = 0ull;
        SC(cudaHostGetDevicePointer(reinterpret_cast<void **>(&s->d_abort_host), s->h_abort, 0));
        const char *e = getenv("DAISY_EXCHANGE_TIMEOUT_MS");
        if (e && atof(e) > 0.0) s->timeout_ns = (unsigned long long)(atof(e) * 1e6);
    }
    SC(cudaMalloc(&s->d_flags, sizeof(unsigned long long) * 17));
    SC(cudaMemset(s->d_flags, 0, sizeof(unsigned long long) * 17));
    SC(cudaMemset(s->d_res[0], 0, exb));
    SC(cudaMemset(s->d_res[1], 0, exb));
    SC(cudaEventCreate(&s->e0));
    SC(cudaEventCreate(&s->e1));
    if (s->use_mma) {
        s->ncolsP = ((s->G * s->n + 63) / 64) * 64;
        SC(cudaMalloc(&s->d_split, sizeof(float) * (size_t)2 * s->Kp * s->ncolsP));
    }
    if (s->use_tma) { int mrc = make_maps(s); if (mrc) { free_solver(s); return mrc; } }
    {
        std::vector<float> ex;
        to_exchange(s, E, ex);
        SC(cudaMemcpy(s->d_E, ex.data(), exb, cudaMemcpyHostToDevice));
        // M padded to Kp x Kp (extra bands are inert: zero rows/columns)
        std::vector<float> Mp((size_t)nmat * s->Kp * s->Kp, 0.0f);
        for (int m = 0; m < nmat; m++)
            for (int j = 0; j < K; j++)
                for (int i = 0; i < K; i++) Mp[((size_t)m * s->Kp + j) * s->Kp + i] = M[((size_t)m * K + j) * K + i];
        SC(cudaMemcpy(s->d_M, Mp.data(), sizeof(float) * Mp.size(), cudaMemcpyHostToDevice));
        if (s->nloc > 0) SC(cudaMemcpy(s->d_mat, mat_idx + s->row0, sizeof(int) * (size_t)s->nloc, cudaMemcpyHostToDevice));
    }
#undef SC
    // the set-up above used synchronous copies and legacy-stream memsets; the solver's kernels run on ctx->stream, which may be
    // a non-blocking stream (daisy_ctx_set_stream): make sure everything has landed before anything is enqueued there
    { cudaError_t e_ = cudaDeviceSynchronize(); if (e_ != cudaSuccess) { daisy_set_error("daisy_solver_create: %s", cudaGetErrorString(e_)); free_solver(s); return DAISY_E_CUDA; } }
    compute_sums_from_host(s, E);
    s->e_sums = s->sums;
    *out = s;
    rc = daisy_solver_reset(s);
    if (rc) { free_solver(s); *out = nullptr; }
    return rc;
}

extern "C" void daisy_solver_destroy(daisy_solver *s) {
    if (s && s->ctx) cudaSetDevice(s->ctx->device);
    free_solver(s);
}

extern "C" int daisy_solver_reset(daisy_solver *s) {
    DZ_REQUIRE(s, DAISY_E_INVALID, "daisy_solver_reset: null solver");
    DZ_CUDA(cudaSetDevice(s->ctx->device));
    cudaStream_t st = s->ctx->stream;
    size_t exb = sizeof(float) * (size_t)s->G * s->bstride;
    // fused exchange: a slower rank may still be storing its block of the last pass into this rank's buffers
    if (s->fused && s->seq > 0) {
        k_wait_flags<<<1, 32, 0, st>>>(s->d_flags, s->G, s->seq, s->timeout_ns, s->d_abort_host);
        DZ_CUDA(cudaGetLastError());
    }
    // a pass that gave up on a peer leaves its counters half way and the abort markers set: start clean
    DZ_CUDA(cudaMemsetAsync(s->d_done, 0, sizeof(unsigned int), st));
    DZ_CUDA(cudaMemsetAsync(s->d_rb_arrive, 0, sizeof(unsigned int) * (size_t)s->nrb_tma, st));
    DZ_CUDA(cudaMemsetAsync(s->d_flags + 16, 0, sizeof(unsigned long long), st));
    if (s->h_abort) *s->h_abort = 0ull;
    // residualvector = emission; lightningvalues = emission                                 Lightning.h:159-165
    s->cur = 0;
    DZ_CUDA(cudaMemcpyAsync(s->d_res[0], s->d_E, exb, cudaMemcpyDeviceToDevice, st));
    DZ_CUDA(cudaMemcpyAsync(s->d_B, s->d_E + (size_t)s->rank * s->bstride, sizeof(float) * (size_t)s->Kp * s->n, cudaMemcpyDeviceToDevice, st));
    DZ_CUDA(cudaStreamSynchronize(st));
    s->numpasses = 0;
    s->sums = s->e_sums; // band sums of E, totalled once at create
    s->sums_valid = true;
    return DAISY_OK;
}

extern "C" int daisy_solver_step_local(daisy_solver *s) {
    DZ_REQUIRE(s, DAISY_E_INVALID, "daisy_solver_step_local: null solver");
    DzRange range_("gather pass");
    DZ_CUDA(cudaSetDevice(s->ctx->device));
    cudaStream_t st = s->ctx->stream;
    if (!s->chained) DZ_CUDA(cudaEventRecord(s->e0, st)); // an event between two passes would keep them from overlapping
    int rc = launch_pass(s);
    if (rc) return rc;
    if (!s->chained) DZ_CUDA(cudaEventRecord(s->e1, st));
    s->events_recorded = !s->chained;
    return DAISY_OK;
}

extern "C" int daisy_solver_set_chained(daisy_solver *s, int on) {
    DZ_REQUIRE(s, DAISY_E_INVALID, "daisy_solver_set_chained: null solver");
    s->chained = on != 0;
    return DAISY_OK;
}

extern "C" int daisy_solver_exchange_info(daisy_solver *s, void **d_next_buffer, int64_t *block_bytes, int64_t *total_bytes) {
    DZ_REQUIRE(s, DAISY_E_INVALID, "daisy_solver_exchange_info: null solver");
    if (d_next_buffer) *d_next_buffer = s->d_res[s->cur ^ 1];
    if (block_bytes) *block_bytes = (int64_t)sizeof(float) * s->bstride;
    if (total_bytes) *total_bytes = (int64_t)sizeof(float) * s->bstride * s->G;
    return DAISY_OK;
}

static int fetch_sums(daisy_solver *s) {
    // per-rank partial band sums sit in the tail of every block of the current buffer
    cudaStream_t st = s->ctx->stream;
    unsigned long long timed_out = 0;
    if (s->fused) {
        k_wait_flags<<<1, 32, 0, st>>>(s->d_flags, s->G, s->seq, s->timeout_ns, s->d_abort_host); // every rank's block of the last pass has landed
        DZ_CUDA(cudaGetLastError());
        DZ_CUDA(cudaMemcpyAsync(&timed_out, s->d_flags + 16, sizeof(timed_out), cudaMemcpyDeviceToHost, st));
    }
    std::vector<double> tails((size_t)s->G * s->Kp);
    for (int g = 0; g < s->G; g++)
        DZ_CUDA(cudaMemcpyAsync(&tails[(size_t)g * s->Kp], s->d_res[s->cur] + (size_t)g * s->bstride + s->sums_off,
                                sizeof(double) * s->Kp, cudaMemcpyDeviceToHost, st));
    DZ_CUDA(cudaStreamSynchronize(st));
    if (timed_out) {
        daisy_set_error("fused exchange: a peer's block of pass %llu did not arrive within %.1f s; the solver is stopped until daisy_solver_reset",
                        timed_out, (double)s->timeout_ns * 1e-9);
        return DAISY_E_STATE;
    }
    float ms = 0.f;
    if (s->events_recorded) {
        if (cudaEventElapsedTime(&ms, s->e0, s->e1) == cudaSuccess) s->last_ms = ms;
        else cudaGetLastError(); // not an error of this call's work: do not leave it for the next cudaGetLastError()
    }
    s->sums.assign(s->K, 0.0);
    for (int k = 0; k < s->K; k++) {
        double v = 0.0;
        for (int g = 0; g < s->G; g++) v += tails[(size_t)g * s->Kp + k];
        s->sums[k] = v;
    }
    s->sums_valid = true;
    return DAISY_OK;
}

// band_sums == NULL: nothing is read back and the call does not wait for the device (back-to-back passes);
// the sums are fetched lazily by daisy_solver_band_sums / daisy_solver_converge.
extern "C" int daisy_solver_step_finish(daisy_solver *s, double *band_sums) {
    DZ_REQUIRE(s, DAISY_E_INVALID, "daisy_solver_step_finish: null solver");
    DZ_CUDA(cudaSetDevice(s->ctx->device));
    s->cur ^= 1;
    s->numpasses++;
    s->sums_valid = false;
    if (!band_sums) return DAISY_OK;
    int rc = fetch_sums(s);
    if (rc) return rc;
    memcpy(band_sums, s->sums.data(), sizeof(double) * s->K);
    return DAISY_OK;
}

extern "C" int daisy_solver_step(daisy_solver *s, double *band_sums) {
    DZ_REQUIRE(s, DAISY_E_INVALID, "daisy_solver_step: null solver");
    DZ_REQUIRE(s->G == 1, DAISY_E_STATE, "daisy_solver_step: partitioned solver needs step_local / exchange / step_finish");
    int rc = daisy_solver_step_local(s);
    if (rc) return rc;
    return daisy_solver_step_finish(s, band_sums);
}

// ---- fused exchange over peer memory (multi-GPU, one process per GPU) ------------------------------------------
extern "C" int daisy_solver_ipc_handles(daisy_solver *s, void *handles192) {
    DZ_REQUIRE(s && handles192, DAISY_E_INVALID, "daisy_solver_ipc_handles: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DZ_CUDA(cudaSetDevice(s->ctx->device));
    cudaIpcMemHandle_t h[3];
    DZ_CUDA(cudaIpcGetMemHandle(&h[0], s->d_res[0]));
    DZ_CUDA(cudaIpcGetMemHandle(&h[1], s->d_res[1]));
    DZ_CUDA(cudaIpcGetMemHandle(&h[2], s->d_flags));
    memcpy(handles192, h, sizeof(h));
    return DAISY_OK;
}

extern "C" int daisy_solver_set_peers(daisy_solver *s, const void *handles, int nranks) {
    DZ_REQUIRE(s && handles, DAISY_E_INVALID, "daisy_solver_set_peers: null argument");
    DZ_REQUIRE(nranks == s->G && nranks <= 16, DAISY_E_INVALID, "daisy_solver_set_peers: nranks must match the partition (<= 16)");
    DZ_REQUIRE(!s->fused, DAISY_E_STATE, "daisy_solver_set_peers: peers already set");
    DZ_CUDA(cudaSetDevice(s->ctx->device));
    for (int g = 0; g < nranks; g++) {
        if (g == s->rank) { s->peer_res[0][g] = s->d_res[0]; s->peer_res[1][g] = s->d_res[1]; s->peer_flags[g] = s->d_flags; continue; }
        cudaIpcMemHandle_t h[3];
        memcpy(h, (const char *)handles + (size_t)g * sizeof(h), sizeof(h));
        void *p0 = nullptr, *p1 = nullptr, *p2 = nullptr;
        DZ_CUDA(cudaIpcOpenMemHandle(&p0, h[0], cudaIpcMemLazyEnablePeerAccess));
        DZ_CUDA(cudaIpcOpenMemHandle(&p1, h[1], cudaIpcMemLazyEnablePeerAccess));
        DZ_CUDA(cudaIpcOpenMemHandle(&p2, h[2], cudaIpcMemLazyEnablePeerAccess));
        s->peer_res[0][g] = (float *)p0; s->peer_res[1][g] = (float *)p1; s->peer_flags[g] = (unsigned long long *)p2;
    }
    s->fused = true;
    s->peers_ipc = true;
    return DAISY_OK;
}

int dz_solver_get_buffers(daisy_solver *s, float **res0, float **res1, unsigned long long **flags) {
    *res0 = s->d_res[0]; *res1 = s->d_res[1]; *flags = s->d_flags;
    return DAISY_OK;
}

int dz_solver_set_peer_pointers(daisy_solver *s, float *const *res0, float *const *res1, unsigned long long *const *flags, int nranks) {
    DZ_REQUIRE(s && res0 && res1 && flags, DAISY_E_INVALID, "solver_set_peer_pointers: null argument");
    DZ_REQUIRE(nranks == s->G && nranks <= 16, DAISY_E_INVALID, "solver_set_peer_pointers: nranks must match the partition (<= 16)");
    DZ_REQUIRE(!s->fused, DAISY_E_STATE, "solver_set_peer_pointers: peers already set");
    for (int g = 0; g < nranks; g++) { s->peer_res[0][g] = res0[g]; s->peer_res[1][g] = res1[g]; s->peer_flags[g] = flags[g]; }
    s->fused = true;
    s->peers_ipc = false;
    return DAISY_OK;
}

// ---- slice upload for the fused exchange: every rank uploads ONLY its own rows of B and of the residual; the residual slice
// (with its band sums) is then handed to every rank over NVLink exactly like the output of a pass -- block `rank` of every
// rank's next buffer, flag[rank] = new pass number -- so the next pass finds the whole vector in place.
__global__ void k_raise_flags(int npeers, unsigned long long seq, const unsigned long long *abort_marker, unsigned long long *f0, unsigned long long *f1,
                              unsigned long long *f2, unsigned long long *f3, unsigned long long *f4, unsigned long long *f5, unsigned long long *f6,
                              unsigned long long *f7, unsigned long long *f8, unsigned long long *f9, unsigned long long *f10, unsigned long long *f11,
                              unsigned long long *f12, unsigned long long *f13, unsigned long long *f14, unsigned long long *f15) {
    unsigned long long *f[16] = { f0, f1, f2, f3, f4, f5, f6, f7, f8, f9, f10, f11, f12, f13, f14, f15 };
    const int g = threadIdx.x;
    if (g >= npeers || *reinterpret_cast<const volatile unsigned long long *>(abort_marker) != 0ull) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f[g]), "l"(seq) : "memory");
}

extern "C" int daisy_solver_write_slices(daisy_solver *s, const float *B_local, const float *residual_local) {
    DZ_REQUIRE(s && B_local && residual_local, DAISY_E_INVALID, "daisy_solver_write_slices: null argument");
    DZ_REQUIRE(s->G == 1 || s->fused, DAISY_E_STATE, "daisy_solver_write_slices: a partitioned solver needs the fused exchange (daisy_solver_set_peers)");
    DZ_CUDA(cudaSetDevice(s->ctx->device));
    cudaStream_t st = s->ctx->stream;
    if (s->G == 1) return daisy_solver_write(s, B_local, residual_local);
    // every rank has finished the last pass (and with it all reads of the buffer about to be overwritten)
    if (s->seq > 0) {
        k_wait_flags<<<1, 32, 0, st>>>(s->d_flags, s->G, s->seq, s->timeout_ns, s->d_abort_host);
        DZ_CUDA(cudaGetLastError());
    }
    const int nxt = s->cur ^ 1;
    float *own = s->d_res[nxt] + (size_t)s->rank * s->bstride;
    if (s->nloc > 0) {
        DZ_CUDA(cudaMemcpy2DAsync(s->d_B, sizeof(float) * s->n, B_local, sizeof(float) * s->nloc, sizeof(float) * s->nloc, s->K, cudaMemcpyHostToDevice, st));
        DZ_CUDA(cudaMemcpy2DAsync(own, sizeof(float) * s->n, residual_local, sizeof(float) * s->nloc, sizeof(float) * s->nloc, s->K, cudaMemcpyHostToDevice, st));
    }
    std::vector<double> tail((size_t)s->Kp, 0.0);
    for (int k = 0; k < s->K; k++) tail[(size_t)k] = host_sum(residual_local + (size_t)k * s->nloc, (size_t)s->nloc);
    DZ_CUDA(cudaMemcpyAsync(own + s->sums_off, tail.data(), sizeof(double) * s->Kp, cudaMemcpyHostToDevice, st));
    DZ_CUDA(cudaStreamSynchronize(st)); // `tail` is pageable host memory
    for (int g = 0; g < s->G; g++) {
        if (g == s->rank) continue;
        DZ_CUDA(cudaMemcpyAsync(s->peer_res[nxt][g] + (size_t)s->rank * s->bstride, own, sizeof(float) * (size_t)s->bstride, cudaMemcpyDeviceToDevice, st));
    }
    s->seq++;
    unsigned long long *f[16];
    for (int g = 0; g < 16; g++) f[g] = (g < s->G) ? s->peer_flags[g] + s->rank : nullptr;
    k_raise_flags<<<1, 32, 0, st>>>(s->G, s->seq, s->d_flags + 16, f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7], f[8], f[9], f[10], f[11], f[12],
                                    f[13], f[14], f[15]);
    DZ_CUDA(cudaGetLastError());
    s->cur = nxt;
    s->sums_valid = false;
    return DAISY_OK;
}

// one pass with the exchange fused into the epilogue kernel: wait for the blocks of the previous pass, gather, store
// this rank's new block into every rank's next buffer and raise the flags.  No host synchronisation, no NCCL call.
extern "C" int daisy_solver_step_fused(daisy_solver *s, double *band_sums) {
    DZ_REQUIRE(s, DAISY_E_INVALID, "daisy_solver_step_fused: null solver");
    DZ_REQUIRE(s->fused, DAISY_E_STATE, "daisy_solver_step_fused: call daisy_solver_set_peers first");
    DzRange range_("gather pass + fused exchange");
    DZ_CUDA(cudaSetDevice(s->ctx->device));
    cudaStream_t st = s->ctx->stream;
    if (s->h_abort && *reinterpret_cast<volatile unsigned long long *>(s->h_abort) != 0ull) {
        daisy_set_error("fused exchange: a peer's block of pass %llu did not arrive within %.1f s; the solver is stopped until daisy_solver_reset",
                        *s->h_abort, (double)s->timeout_ns * 1e-9);
        return DAISY_E_STATE;
    }
    DZ_CUDA(cudaEventRecord(s->e0, st));
    const unsigned long long wait_seq = s->seq; // the blocks of the previous pass (0: first pass, nothing to wait for)
    s->seq++;
    int rc = launch_pass(s, wait_seq);
    if (rc) return rc;
    DZ_CUDA(cudaEventRecord(s->e1, st));
    s->events_recorded = true;
    return daisy_solver_step_finish(s, band_sums);
}

static bool unconverged(const daisy_solver *s, double threshold, int per_band) {
    if (per_band) {
        for (int k = 0; k < s->K; k++) if (s->sums[k] > threshold) return true;
        return false;
    }
    double tot = 0.0;
    for (int k = 0; k < s->K; k++) tot += s->sums[k];
    return tot > threshold;
}

extern "C" int daisy_solver_converge(daisy_solver *s, double threshold, int per_band, int max_passes, int *passes_out) {
    DZ_REQUIRE(s, DAISY_E_INVALID, "daisy_solver_converge: null solver");
    DZ_REQUIRE(s->G == 1, DAISY_E_STATE, "daisy_solver_converge: partitioned solver is driven by the host exchange loop (or daisy_group_solver_converge)");
    DzRange range_("daisy_solver_converge");
    int done = 0;
    std::vector<double> tmp(s->K);
    if (!s->sums_valid) { int rc = fetch_sums(s); if (rc) return rc; }
    while (unconverged(s, threshold, per_band) && (max_passes <= 0 || done < max_passes)) {
        int rc = daisy_solver_step(s, tmp.data());
        if (rc) return rc;
        done++;
    }
    if (passes_out) *passes_out = s->numpasses;
    return DAISY_OK;
}

extern "C" int daisy_solver_numpasses(daisy_solver *s) { return s ? s->numpasses : DAISY_E_INVALID; }

extern "C" int daisy_solver_band_sums(daisy_solver *s, double *band_sums) {
    DZ_REQUIRE(s && band_sums, DAISY_E_INVALID, "daisy_solver_band_sums: null argument");
    if (!s->sums_valid) { DZ_CUDA(cudaSetDevice(s->ctx->device)); int rc = fetch_sums(s); if (rc) return rc; }
    memcpy(band_sums, s->sums.data(), sizeof(double) * s->K);
    return DAISY_OK;
}

extern "C" int daisy_solver_read(daisy_solver *s, float *B, float *residual) {
    DZ_REQUIRE(s, DAISY_E_INVALID, "daisy_solver_read: null solver");
    DZ_CUDA(cudaSetDevice(s->ctx->device));
    // local rows only: out[k*nloc + pl]; both copies ride the solver's stream behind the passes, one synchronisation at the end
    cudaStream_t st = s->ctx->stream;
    if (B)
        DZ_CUDA(cudaMemcpy2DAsync(B, sizeof(float) * s->nloc, s->d_B, sizeof(float) * s->n, sizeof(float) * s->nloc, s->K, cudaMemcpyDeviceToHost, st));
    if (residual)
        DZ_CUDA(cudaMemcpy2DAsync(residual, sizeof(float) * s->nloc, s->d_res[s->cur] + (size_t)s->rank * s->bstride, sizeof(float) * s->n,
                                  sizeof(float) * s->nloc, s->K, cudaMemcpyDeviceToHost, st));
    DZ_CUDA(cudaStreamSynchronize(st));
    return DAISY_OK;
}

extern "C" int daisy_solver_write(daisy_solver *s, const float *B, const float *residual) {
    DZ_REQUIRE(s && B && residual, DAISY_E_INVALID, "daisy_solver_write: null argument");
    DZ_REQUIRE(s->G == 1, DAISY_E_STATE, "daisy_solver_write: single-GPU solvers only");
    DZ_CUDA(cudaSetDevice(s->ctx->device));
    cudaStream_t st = s->ctx->stream;
    DZ_CUDA(cudaMemcpy2DAsync(s->d_B, sizeof(float) * s->n, B, sizeof(float) * s->N, sizeof(float) * s->N, s->K, cudaMemcpyHostToDevice, st));
    DZ_CUDA(cudaMemcpy2DAsync(s->d_res[s->cur], sizeof(float) * s->n, residual, sizeof(float) * s->N, sizeof(float) * s->N, s->K,
                              cudaMemcpyHostToDevice, st));
    compute_sums_from_host(s, residual); // while the copies are in flight
    DZ_CUDA(cudaStreamSynchronize(st));
    s->sums_valid = true;
    return DAISY_OK;
}

// partitioned counterpart of daisy_solver_write: B_local is K x nloc (this rank's rows), residual_full is K x N (every
// rank holds the whole residual vector, as after an exchange)
extern "C" int daisy_solver_write_partitioned(daisy_solver *s, const float *B_local, const float *residual_full) {
    DZ_REQUIRE(s && B_local && residual_full, DAISY_E_INVALID, "daisy_solver_write_partitioned: null argument");
    DZ_CUDA(cudaSetDevice(s->ctx->device));
    cudaStream_t st = s->ctx->stream;
    if (s->fused && s->seq > 0) { // peers' stores of the last pass have landed
        k_wait_flags<<<1, 32, 0, st>>>(s->d_flags, s->G, s->seq, s->timeout_ns, s->d_abort_host);
        DZ_CUDA(cudaGetLastError());
    }
    if (s->nloc > 0)
        DZ_CUDA(cudaMemcpy2DAsync(s->d_B, sizeof(float) * s->n, B_local, sizeof(float) * s->nloc, sizeof(float) * s->nloc, s->K, cudaMemcpyHostToDevice, st));
    for (int g = 0; g < s->G; g++) {
        const int c0 = g * s->n, w = (c0 < s->N) ? ((s->N - c0 < s->n) ? s->N - c0 : s->n) : 0;
        if (w > 0)
            DZ_CUDA(cudaMemcpy2DAsync(s->d_res[s->cur] + (size_t)g * s->bstride, sizeof(float) * s->n, residual_full + c0, sizeof(float) * s->N,
                                      sizeof(float) * w, s->K, cudaMemcpyHostToDevice, st));
    }
    compute_sums_from_host(s, residual_full); // while the copies are in flight
    DZ_CUDA(cudaStreamSynchronize(st));
    s->sums_valid = true;
    return DAISY_OK;
}

// kernels one pass launches in this solver's configuration (the benchmark reports it)
extern "C" int daisy_solver_launches_per_pass(daisy_solver *s) {
    if (!s) return DAISY_E_INVALID;
    if (s->fused_epi) return 1;                         // k_gather_tma with the epilogue and the exchange wait inside
    int n = 2;                                          // streaming kernel + k_gather_epilogue
    if (s->use_mma) n += 1;                             // k_split_residual
    if (s->fused) n += 1;                               // k_wait_flags
    return n;
}

extern "C" int daisy_solver_last_step_ms(daisy_solver *s, double *ms) {
    DZ_REQUIRE(s && ms, DAISY_E_INVALID, "daisy_solver_last_step_ms: null argument");
    *ms = s->last_ms;
    return DAISY_OK;
}
