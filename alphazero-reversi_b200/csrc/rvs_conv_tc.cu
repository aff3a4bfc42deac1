// rvs_conv_tc.cu -- 3x3 convolution C->C (+ folded BN bias, optional residual, ReLU) as an
// implicit GEMM on the 5th-generation tensor cores: TMA -> shared memory -> tcgen05.mma -> TMEM ->
// tcgen05.ld epilogue.  sm_100a only.
//
// GEMM view per CTA PAIR (cta_group::2):  D[256 pixels, C couts] += A_tap[256 pixels, cin] * W_tap[C couts, cin]^T
//   M = 2 x 128: each CTA supplies one "tile" = two boards, rows ordered (y, board, x): with this order a
//       vertical tap shift dy is a shift by 16 rows = 2048 B, i.e. a whole number of 1024-byte swizzle atoms.
//   N = C couts; each CTA holds HALF of the weight rows (resident for C <= 128, streamed for C = 256).
//   K = 9 taps x cin, consumed as 3 (dx) x cin/64 stages; each stage is ONE TMA box of the
//       activations: 10 rows of y (halo -1..8) x 2 boards x 8 x (shifted by dx) x 64 channels,
//       out-of-bounds rows/columns zero-filled by the TMA unit = the convolution's zero padding.
//       The three vertical taps of that dx reuse the same box at row offsets 0 / 16 / 32.
// Activations live in HBM in the same tiled order [tile][y][board][x][c] (rvs_conv_tc.cuh:
// act_row), so accumulator row r of a tile is simply row r of the output tile.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (leader CTA only),
// warps 2-5 = epilogue (TMEM lane quarter = warp_id % 4).  Two accumulators in TMEM ping-pong so that the
// epilogue of tile i overlaps the MMAs of tile i+1.  (The first-generation 1-CTA kernel -- N = 64 per CTA,
// tensor pipe at 31 % of its throughput, profiles/superseded/ncu_conv_tc1_r1.txt -- was removed in round 2.)
#include "rvs_conv_tc.cuh"

#include <cuda.h>
#include <stdlib.h>

#include "rvs_common.cuh"

namespace rvs {

namespace {

constexpr int kTileRows = 128;              // pixels per tile (2 boards)
constexpr int kABytes = 10 * 2 * 8 * 128;   // one activation stage of the streamed (256-filter) kernel: 160 rows x 128 B
// Resident-weight kernels (C <= 128): ONE activation box per 64-channel chunk serves all nine taps.  The box is
// 10 rows of y (halo -1..8) x 2 boards x 16 x-slots (x = -1..14; x < 0 and x > 7 are zero-filled by the TMA unit)
// x 64 channels, so a (y, board) line is exactly two 1024-byte swizzle atoms and a horizontal tap shift dx is a
// start-address offset of dx rows (128 B) INSIDE the atom (make_desc_x).  Compared with one shifted box per dx this fetches every input element once instead of three
// times (L2 -> SM traffic per tile 40 KB instead of 120 KB per chunk pair) and writes 80 KB instead of 120 KB of
// shared memory per tile -- the 128-filter layer was bound by exactly these two.
#ifndef RVS_CONV_XSLOTS
#define RVS_CONV_XSLOTS 16
#endif
constexpr int kXSlots = RVS_CONV_XSLOTS;
constexpr int kABytesX = 10 * 2 * kXSlots * 128;  // 40 KB
constexpr int kThreads = 192;   // streamed 256-filter kernel: warp 0 producer, warp 1 MMA issuer, warps 2-5 epilogue

struct Impl {
    CUtensorMap w_map2;   // box of C/2 weight rows (one CTA's half of a tap)
    int cin = 0;
    // TMA descriptors of the activation buffers this layer has been launched on, keyed by (base pointer, capacity):
    // 3 ping-pong buffers x 2 half-batches of a pipelined search
    static constexpr int kActSlots = 8;
    const void* act_ptr[kActSlots] = {};
    int64_t act_cap[kActSlots] = {};
    CUtensorMap act_map[kActSlots];
    int n_act = 0;
};

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeFn)p;
    }
    return fn;
}

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
// 256-bit global accesses (LDG/STG.E.ENL2.256 on sm_100a): one full 32-byte sector per instruction
__device__ __forceinline__ void ldg256(const void* p, uint32_t* v) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* v) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (=1, unused for swizzled K-major)
//   [32,46) stride byte offset >> 4 (1024 B between 8-row groups) | [46,48) version = 1 (sm_100)
//   [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// the same with 2048 B between 8-row groups (16 x-slots per (y, board) line).  The start address may sit 1 or 2 rows
// (128 / 256 B) into a swizzle atom (horizontal tap shift): the tensor core applies the 128-byte swizzle to the
// ABSOLUTE shared-memory address bits, exactly like the TMA unit that wrote the box, so no descriptor field has to
// describe the phase.  Measured on B200: with the descriptor's base-offset field [49,52) set to the row phase the
// results are wrong, with 0 all network goldens pass (tests/test_gpu_net.py).
__device__ __forceinline__ uint64_t make_desc_x(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)((kXSlots * 128) >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16
// [10,13)=1, A/B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)  -> Cfg2::IDESC / Cfg2S::IDESC

// =============================================================================================
// CTA pairs (cta_group::2).  Profiling a 1-CTA kernel showed the tensor pipe busy 73 % of
// the time at only 31 % of its throughput: with N = 64 every MMA streams 4 KB of A + 2 KB of B
// from shared memory for 32 cycles of math (192 B/cycle > the 128 B/cycle SMEM port).  Here a CTA
// pair computes D[256 px, C couts]: each CTA supplies its own 128-pixel tile (A) and HALF of the
// weight rows (B, still resident), so one MMA reads 4 KB + 2 KB per CTA for 64 cycles of math
// (96 B/cycle) and every activation byte is fetched once per cout instead of twice.
//   * both CTAs run a TMA producer (own A tile, own W half) that signals the LEADER's barriers
//     (cp.async.bulk.tensor ... .cta_group::2 with the peer bit of the barrier address cleared);
//   * only the leader (cluster rank 0) issues tcgen05.mma.cta_group::2 and commits with
//     .multicast::cluster so that "stage free" / "accumulator ready" arrive in both CTAs;
//   * both CTAs run the epilogue on their own TMEM (128 lanes x C columns, double buffered); the
//     peer's epilogue warps arrive remotely on the leader's "accumulator free" barrier.
// =============================================================================================
// programmatic dependent launch: the next layer's prologue (barrier init, TMEM alloc, 144 KB of
// weights) overlaps the tail of this layer; activations are only touched after pdl_wait()
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // cute::Sm100MmaPeerBitMask: address of the same offset in the even CTA
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* map, uint32_t leader_bar, uint32_t dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(leader_bar & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma2_load_5d(const CUtensorMap* map, uint32_t leader_bar, uint32_t dst, int c0, int c1, int c2,
                                             int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
        "l"(map), "r"(leader_bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tc2_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tc2_commit_mc(uint32_t bar) {  // arrive on `bar` in BOTH CTAs of the pair
    asm volatile(
        "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(bar)
        : "memory");
}
// Barrier traffic inside the pair carries NO generic-proxy memory dependency: operands travel TMA -> shared memory ->
// tensor core (async proxy, ordered by complete_tx / tcgen05.commit) and results tensor core -> TMEM -> registers
// (ordered by tcgen05.fence / tcgen05.wait::ld).  So the remote arrive and the waits use the default CTA-scope
// semantics, as CUTLASS's ClusterBarrier does.  The `.release.cluster` / `.acquire.cluster` forms used in round 1
// compiled to MEMBAR.ALL.GPU + ERRBAR per arrive and to CCTL.IVALL (invalidate all of L1) per successful wait: a
// GPU-wide fence per epilogue warp per tile that waited for the tile's 32 KB of output stores to drain (ncu:
// stall_membar 17 % of the epilogue warps' samples).
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {  // arrive on the leader CTA's barrier from either CTA
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}

// HEAD variant (last tower layer): the policy / value 1x1 convolutions + BN + ReLU
// (network.py:104-105, 111-112) are three dot products over the output row the epilogue already
// holds in registers, so the 67 MB activation write and its re-read by the head kernel disappear;
// weights travel in the kernel parameter space (constant cache, warp-uniform reads).
using HeadW = ConvHeadW;  // rvs_conv_tc.cuh

#ifdef RVS_CONV_PROBE  // debug variant only (tools/probe_conv.py): where the roles of the 128-filter kernel wait
__device__ long long g_conv_probe[148 * 16];
#define PROBE_T0() const long long _pt = clock64()
#define PROBE_ADD(x) (x) += clock64() - _pt
#else
#define PROBE_T0()
#define PROBE_ADD(x)
#endif

template <int C, int CIN>
struct Cfg2 {
    static constexpr int KC = CIN / 64;
    static constexpr int NH = C / 2;                        // weight rows (couts) held by each CTA
    static constexpr int W_TILE = NH * 128;                 // bytes of one (tap, kc) weight tile per CTA
    static constexpr int W_TILES = 9 * KC;
    static constexpr int STAGES = (C == 128 && CIN == 128) ? (kXSlots <= 10 ? 3 : 2) : (C == 128 ? 3 : 4);  // 128->128: 144 KB weights + 2 x 40 KB = 224 KB
    static constexpr int TMEM_COLS = 2 * C;                 // two accumulators of C fp32 columns
    // K = 16 steps per 64-channel chunk.  The 64 -> 128 instantiation is the network's FIRST layer: only
    // channels 0..2 of its input tiles are non-zero, so the steps over channels 16..63 would multiply zeros
    static constexpr int KSTEPS = (C == 128 && CIN == 64) ? 1 : 4;
    static constexpr int EW = 4;                                // epilogue warps (4 or 8, see the epilogue's comment)
    static constexpr int THREADS = 64 + 32 * EW;                // warp 0 TMA producer, warp 1 MMA issuer, then the epilogue
    static constexpr int SMEM = W_TILES * W_TILE + STAGES * kABytesX + 1024 /*align*/ + 1280 /*barriers, bias*/;
    static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
};

template <int C, int CIN, bool HEAD, bool FRESH>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg2<C, CIN>::THREADS, 1)
conv3x3_tc2_kernel(const __grid_constant__ CUtensorMap a_map, const __grid_constant__ CUtensorMap w_map,
                   const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ out,
                   const float* __restrict__ bias, int n_tiles_arg, const __grid_constant__ HeadW head,
                   float* __restrict__ feat, const int* __restrict__ n_boards_dev, int rev) {
    using K = Cfg2<C, CIN>;
    // Compacted leaf batches: the number of boards is only known on the device.  Normally it was written at least two
    // kernels upstream and is read here at once (visible even when this launch overlaps its predecessor's tail).  FRESH:
    // the kernel just before this one wrote it (the first layer after the tree step of a wave-1 search, which releases
    // its dependents early), so every role re-reads it after its griddepcontrol.wait.  That is a template parameter and
    // not a flag because a trip count that comes from a volatile load is no longer a UNIFORM value for the compiler: the
    // MMA issuer's descriptor arithmetic then leaves the uniform datapath (5 R2UR per tcgen05.mma) and a tower layer ran
    // 4 % slower; the first layer issues 9 MMAs per tile and does not care.
    int n_tiles = (n_boards_dev && !FRESH) ? (*n_boards_dev + 1) >> 1 : n_tiles_arg;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t w_s = base;
    const uint32_t a_s = base + K::W_TILES * K::W_TILE;
    unsigned char* tail = gen + K::W_TILES * K::W_TILE + K::STAGES * kABytesX;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 192);
    float* sbias = reinterpret_cast<float*>(tail + 256);  // [C] <= 256 floats
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };                       // used in the leader
    auto EMPTY = [&](int s) { return bar0 + 8u * (K::STAGES + s); };        // both CTAs (multicast commit)
    const uint32_t WFULL = bar0 + 8u * (2 * K::STAGES);                     // leader
    auto ACC_FULL = [&](int a) { return bar0 + 8u * (2 * K::STAGES + 1 + a); };   // both CTAs
    auto ACC_EMPTY = [&](int a) { return bar0 + 8u * (2 * K::STAGES + 3 + a); };  // leader, one arrival per epilogue warp of both CTAs

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    int n_iters = (n_tiles + 2 * n_pairs - 1) / (2 * n_pairs);  // same trip count in both CTAs of a pair
    auto wait_for_inputs = [&]() {
        pdl_wait();
        if (FRESH && n_boards_dev) {
            n_tiles = (__ldcg(n_boards_dev) + 1) >> 1;
            n_iters = (n_tiles + 2 * n_pairs - 1) / (2 * n_pairs);
        }
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < K::STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        mbar_init(WFULL, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(ACC_FULL(a), 1); mbar_init(ACC_EMPTY(a), 2 * K::EW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < C; i += K::THREADS) sbias[i] = bias[i];
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(K::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // barriers of both CTAs are initialised before anyone signals across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) pdl_launch_dependents();

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer (both CTAs) =====
            if (rank == 0) mbar_expect_tx(WFULL, 2 * K::W_TILES * K::W_TILE);  // both halves report to the leader
            for (int tap = 0; tap < 9; ++tap)
                for (int kc = 0; kc < K::KC; ++kc)
                    tma2_load_2d(&w_map, WFULL, w_s + (tap * K::KC + kc) * K::W_TILE, kc * 64, tap * C + (int)rank * K::NH);
            wait_for_inputs();  // weights are constants; activations come from the previous layer
            int stage = 0, phase = 0;
            long long p_empty = 0; (void)p_empty;
            for (int it = 0; it < n_iters; ++it) {
                const int tile = ((rev ? n_iters - 1 - it : it) * n_pairs + pair) * 2 + (int)rank;  // may be >= n_tiles: TMA zero-fills
                for (int kc = 0; kc < K::KC; ++kc) {
                    { PROBE_T0(); mbar_wait_cluster(EMPTY(stage), phase ^ 1); PROBE_ADD(p_empty); }
                    if (rank == 0) mbar_expect_tx(FULL(stage), 2 * kABytesX);
                    tma2_load_5d(&a_map, FULL(stage), a_s + stage * kABytesX, kc * 64, -1, 0, -1, tile);
                    if (++stage == K::STAGES) { stage = 0; phase ^= 1; }
                }
            }
#ifdef RVS_CONV_PROBE
            if (!HEAD && CIN == 128) g_conv_probe[blockIdx.x * 16 + 0] = p_empty;
#endif
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {  // ===== MMA issuer (leader only) =====
            long long p_w = 0, p_acc = 0, p_full = 0, p_start = 0; (void)p_w; (void)p_acc; (void)p_full; (void)p_start;
#ifdef RVS_CONV_PROBE
            p_start = clock64();
#endif
            wait_for_inputs();
            { PROBE_T0(); mbar_wait_cluster(WFULL, 0); PROBE_ADD(p_w); }
            int stage = 0, phase = 0;
            for (int it = 0; it < n_iters; ++it) {
                const int acc = it & 1;
                { PROBE_T0(); mbar_wait_cluster(ACC_EMPTY(acc), ((it >> 1) & 1) ^ 1); PROBE_ADD(p_acc); }
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(acc * C);
                uint32_t accum = 0;
                for (int kc = 0; kc < K::KC; ++kc) {
                    { PROBE_T0(); mbar_wait_cluster(FULL(stage), phase); PROBE_ADD(p_full); }
                    tc_fence_after();
                    const uint32_t a0 = a_s + stage * kABytesX;
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const uint32_t wt = w_s + ((dy * 3 + dx) * K::KC + kc) * K::W_TILE;
                            // tap (dy, dx): input pixel (y + dy - 1, x + dx - 1) = box row ((y + dy) * 2 + board) * 16 + x + dx
                            const uint32_t at = a0 + (uint32_t)(dy * 2 * kXSlots + dx) * 128u;
#pragma unroll
                            for (int k = 0; k < K::KSTEPS; ++k) {
                                tc2_mma(d, make_desc_x(at + k * 32), make_desc(wt + k * 32), K::IDESC, accum);
                                accum = 1;
                            }
                        }
                    }
                    tc2_commit_mc(EMPTY(stage));
                    if (++stage == K::STAGES) { stage = 0; phase ^= 1; }
                }
                tc2_commit_mc(ACC_FULL(acc));
            }
#ifdef RVS_CONV_PROBE
            if (!HEAD && CIN == 128) {
            g_conv_probe[blockIdx.x * 16 + 1] = p_w;
            g_conv_probe[blockIdx.x * 16 + 2] = p_acc;
            g_conv_probe[blockIdx.x * 16 + 3] = p_full;
            g_conv_probe[blockIdx.x * 16 + 4] = clock64() - p_start;
            g_conv_probe[blockIdx.x * 16 + 7] = n_iters;
            }
#endif
        }
    } else {  // ===== epilogue (both CTAs, own tile) =====
        // Warp (q, hh) converts TMEM lane quarter q (32 pixel rows; hardware rule: warp id % 4) x channel slice hh.
        // FOUR warps (one slice each, EW = 4): a tile's conversion is ~900 instructions that one warp per scheduler runs at
        // 0.25 IPC, ~3000 cycles (~5000 with a residual) against ~5000 for the tile's 72 MMAs.  Eight warps (EW = 8, two
        // slices) finish it in half the time, but measured on B200 the MMAs themselves then take 81 instead of 70 cycles
        // each and a tower layer gets 3 % SLOWER (the SM is at its power limit under tensor load: what the other pipes
        // gain, the tensor pipe loses); the epilogue-paced first layer did not gain either (24.3 vs 23.7 us).
        constexpr int CH = C / (K::EW / 4);  // channels per epilogue warp
        const int q = warp & 3;
        const int hh = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        long long p_af = 0, p_e0 = 0; (void)p_af; (void)p_e0;
        wait_for_inputs();
#ifdef RVS_CONV_PROBE
        p_e0 = clock64();
#endif
        for (int it = 0; it < n_iters; ++it) {
            const int acc = it & 1;
            const int tile = ((rev ? n_iters - 1 - it : it) * n_pairs + pair) * 2 + (int)rank;
            const size_t off = ((size_t)tile * kTileRows + row) * C;
            const bool live = tile < n_tiles;
            if constexpr (HEAD) {
                // The three head planes are dot products over ALL channels of a row in a fixed order: the hh = 0 warp of each
                // quarter walks the whole row, its hh = 1 partner only releases the accumulator.
                float hd0 = 0.f, hd1 = 0.f, hd2 = 0.f;
                uint32_t res[C / 2];
                if (hh == 0) {
                    if (residual && live) {
#pragma unroll
                        for (int i = 0; i < C / 16; ++i) ldg256(residual + off + i * 16, res + i * 8);
                    } else {
#pragma unroll
                        for (int i = 0; i < C / 2; ++i) res[i] = 0u;
                    }
                }
                { PROBE_T0(); mbar_wait_cluster(ACC_FULL(acc), (it >> 1) & 1); PROBE_ADD(p_af); }
                tc_fence_after();
                if (hh == 0) {
#pragma unroll
                    for (int h = 0; h < C / 64; ++h) {
                        uint32_t v[64];
                        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C + h * 64);
                        tc_ld32(taddr, v);
                        tc_ld32(taddr + 32, v + 32);
                        tc_wait_ld();
                        if (h == C / 64 - 1) {  // all TMEM reads of this accumulator are done
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_leader(ACC_EMPTY(acc));
                        }
                        if (live) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const int col = 2 * i;
                                const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&res[h * 32 + i]));
                                const float f0 = __uint_as_float(v[col]) + (sbias[h * 64 + col] + t.x);
                                const float f1 = __uint_as_float(v[col + 1]) + (sbias[h * 64 + col + 1] + t.y);
                                // the heads see the same bf16-rounded activations as the unfused path
                                const float2 a = __bfloat1622float2(__floats2bfloat162_rn(fmaxf(f0, 0.f), fmaxf(f1, 0.f)));
                                const int cc = h * 64 + col;
                                hd0 = fmaf(a.x, head.w[0][cc], fmaf(a.y, head.w[0][cc + 1], hd0));
                                hd1 = fmaf(a.x, head.w[1][cc], fmaf(a.y, head.w[1][cc + 1], hd1));
                                hd2 = fmaf(a.x, head.w[2][cc], fmaf(a.y, head.w[2][cc + 1], hd2));
                            }
                        }
                    }
                    if (live) {  // feat[board][plane*64 + px], plane 0/1 policy, 2 value
                        const int y = row >> 4, b = (row >> 3) & 1, x = row & 7;
                        float* fp = feat + ((size_t)tile * 2 + b) * 192 + y * 8 + x;
                        fp[0] = fmaxf(hd0 + head.b[0], 0.f);
                        fp[64] = fmaxf(hd1 + head.b[1], 0.f);
                        fp[128] = fmaxf(hd2 + head.b[2], 0.f);
                    }
                } else {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(ACC_EMPTY(acc));
                }
            } else {
                // residual slice prefetched into registers BEFORE waiting for the accumulator: its HBM/L2
                // latency hides behind the MMAs of this tile instead of extending the epilogue
                uint32_t res[CH / 2];  // CH bf16 = CH/2 words, as CH/16 256-bit loads
                if (residual && live) {
#pragma unroll
                    for (int i = 0; i < CH / 16; ++i) ldg256(residual + off + hh * CH + i * 16, res + i * 8);
                } else {
#pragma unroll
                    for (int i = 0; i < CH / 2; ++i) res[i] = 0u;
                }
                { PROBE_T0(); mbar_wait_cluster(ACC_FULL(acc), (it >> 1) & 1); PROBE_ADD(p_af); }
                tc_fence_after();
                constexpr int CK = CH < 64 ? CH : 64;  // columns per TMEM staging chunk (v[CK])
#pragma unroll
                for (int h = 0; h < CH / CK; ++h) {
                    uint32_t v[CK];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C + hh * CH + h * CK);
#pragma unroll
                    for (int j = 0; j < CK / 32; ++j) tc_ld32(taddr + j * 32, v + j * 32);
                    tc_wait_ld();
                    if (h == CH / CK - 1) {  // all of this warp's TMEM reads of the accumulator are done
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_leader(ACC_EMPTY(acc));
                    }
                    if (live) {
#pragma unroll
                        for (int c16 = 0; c16 < CK / 16; ++c16) {  // 16 couts = one 32-byte sector of bf16
                            uint32_t o[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int col = c16 * 16 + 2 * i;
                                const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&res[h * (CK / 2) + c16 * 8 + i]));
                                float f0 = __uint_as_float(v[col]), f1 = __uint_as_float(v[col + 1]);
                                if constexpr (!(C == 128 && CIN == 64)) {  // the first layer has its bias in K and no residual
                                    f0 += sbias[hh * CH + h * CK + col] + t.x;
                                    f1 += sbias[hh * CH + h * CK + col + 1] + t.y;
                                }
                                const __nv_bfloat162 ob = __floats2bfloat162_rn(fmaxf(f0, 0.f), fmaxf(f1, 0.f));
                                o[i] = *reinterpret_cast<const uint32_t*>(&ob);
                            }
                            stg256(out + off + hh * CH + h * CK + c16 * 16, o);
                        }
                    }
                }
            }
        }
#ifdef RVS_CONV_PROBE
        if (!HEAD && CIN == 128 && warp == 2 && lane == 0) {
            g_conv_probe[blockIdx.x * 16 + 5] = p_af;
            g_conv_probe[blockIdx.x * 16 + 6] = clock64() - p_e0;
        }
#endif
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's shared memory / TMEM stay valid until the leader's MMAs are done
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(K::TMEM_COLS));
    }
}


// =============================================================================================
// 256-filter variant (BASELINE config 4: 20 blocks x 256 filters): the 1.18 MB of one layer's
// weights cannot stay resident, so they STREAM through the same mbarrier ring as the activations.
// A CTA pair computes D[256 px, 256 couts]; a stage is one (dx, 64-channel chunk): the CTA's
// activation box (20 KB) + its half (128 couts) of the three vertical taps' weight tiles
// (3 x 16 KB), consumed by 12 MMAs of M256 x N256 x K16 (128 tensor cycles each: 8 KB of operands
// per CTA per MMA = 64 B/cycle of shared memory, below the 128 B/cycle port).  Two 256-column
// accumulators fill the 512 TMEM columns, so the epilogue of tile i still overlaps the MMAs of
// tile i+1.  L2 -> SM traffic is 68 KB per 1536 tensor cycles per SM (44 B/cycle, at the measured
// ~42 B/cycle/SM L2 ceiling): this layer shape is L2-bandwidth bound, not tensor bound.
// =============================================================================================
struct Cfg2S {
    static constexpr int C = 256;
    static constexpr int KC = C / 64;
    static constexpr int NH = C / 2;                  // weight rows (couts) held by each CTA
    static constexpr int W_TILE = NH * 128;           // 16 KB: one (tap, kc) weight tile per CTA
    static constexpr int STAGE = kABytes + 3 * W_TILE;  // 68 KB
    static constexpr int STAGES = 3;
    static constexpr int TMEM_COLS = 2 * C;           // 512: the whole tensor memory
    static constexpr int SMEM = STAGES * STAGE + 1024 /*align*/ + 1280 /*barriers, bias*/;
    static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv3x3_tc2s_kernel(const __grid_constant__ CUtensorMap a_map, const __grid_constant__ CUtensorMap w_map,
                    const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ out,
                    const float* __restrict__ bias, int n_tiles_arg, const int* __restrict__ n_boards_dev, int rev) {
    using K = Cfg2S;
    // compacted leaf batches: the number of boards is only known on the device (written at least two kernels upstream:
    // this kernel is never the first after the tree step, so it is visible even when this launch overlaps its predecessor's tail)
    const int n_tiles = n_boards_dev ? (*n_boards_dev + 1) >> 1 : n_tiles_arg;
    constexpr int C = K::C;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));
    unsigned char* tail = gen + K::STAGES * K::STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 192);
    float* sbias = reinterpret_cast<float*>(tail + 256);  // [256]
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };                              // leader
    auto EMPTY = [&](int s) { return bar0 + 8u * (K::STAGES + s); };               // both CTAs (multicast commit)
    auto ACC_FULL = [&](int a) { return bar0 + 8u * (2 * K::STAGES + a); };        // both CTAs
    auto ACC_EMPTY = [&](int a) { return bar0 + 8u * (2 * K::STAGES + 2 + a); };   // leader, 8 arrivals

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int n_iters = (n_tiles + 2 * n_pairs - 1) / (2 * n_pairs);
    auto wait_for_inputs = [&]() { pdl_wait(); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < K::STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(ACC_FULL(a), 1); mbar_init(ACC_EMPTY(a), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < C; i += kThreads) sbias[i] = bias[i];
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(K::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) pdl_launch_dependents();

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer (both CTAs): own activation tile + own half of the weights =====
            wait_for_inputs();
            int stage = 0, phase = 0;
            for (int it = 0; it < n_iters; ++it) {
                const int tile = ((rev ? n_iters - 1 - it : it) * n_pairs + pair) * 2 + (int)rank;  // may be >= n_tiles: TMA zero-fills
                for (int dx = 0; dx < 3; ++dx)
                    for (int kc = 0; kc < K::KC; ++kc) {
                        mbar_wait_cluster(EMPTY(stage), phase ^ 1);
                        if (rank == 0) mbar_expect_tx(FULL(stage), 2 * K::STAGE);
                        const uint32_t st = base + stage * K::STAGE;
                        tma2_load_5d(&a_map, FULL(stage), st, kc * 64, dx - 1, 0, -1, tile);
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy)
                            tma2_load_2d(&w_map, FULL(stage), st + kABytes + dy * K::W_TILE, kc * 64, (dy * 3 + dx) * C + (int)rank * K::NH);
                        if (++stage == K::STAGES) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {  // ===== MMA issuer (leader only) =====
            wait_for_inputs();
            int stage = 0, phase = 0;
            for (int it = 0; it < n_iters; ++it) {
                const int acc = it & 1;
                mbar_wait_cluster(ACC_EMPTY(acc), ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(acc * C);
                uint32_t accum = 0;
                for (int s12 = 0; s12 < 3 * K::KC; ++s12) {
                    mbar_wait_cluster(FULL(stage), phase);
                    tc_fence_after();
                    const uint32_t a0 = base + stage * K::STAGE;
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const uint32_t wt = a0 + kABytes + dy * K::W_TILE;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            tc2_mma(d, make_desc(a0 + dy * 2048 + k * 32), make_desc(wt + k * 32), K::IDESC, accum);
                            accum = 1;
                        }
                    }
                    tc2_commit_mc(EMPTY(stage));
                    if (++stage == K::STAGES) { stage = 0; phase ^= 1; }
                }
                tc2_commit_mc(ACC_FULL(acc));
            }
        }
    } else {  // ===== epilogue (both CTAs, own tile): 4 chunks of 64 couts, residual chunk h+1 in flight during chunk h =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        wait_for_inputs();
        for (int it = 0; it < n_iters; ++it) {
            const int acc = it & 1;
            const int tile = ((rev ? n_iters - 1 - it : it) * n_pairs + pair) * 2 + (int)rank;
            const size_t off = ((size_t)tile * kTileRows + row) * C;
            const bool live = tile < n_tiles;
            const bool has_res = residual != nullptr && live;
            uint32_t res[2][32];
#pragma unroll
            for (int i = 0; i < 32; ++i) res[0][i] = 0u;
            if (has_res) {
#pragma unroll
                for (int i = 0; i < 4; ++i) ldg256(residual + off + i * 16, res[0] + i * 8);
            }
            mbar_wait_cluster(ACC_FULL(acc), (it >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                if (h < 3) {
                    if (has_res) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) ldg256(residual + off + (h + 1) * 64 + i * 16, res[(h + 1) & 1] + i * 8);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) res[(h + 1) & 1][i] = 0u;
                    }
                }
                uint32_t v[64];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C + h * 64);
                tc_ld32(taddr, v);
                tc_ld32(taddr + 32, v + 32);
                tc_wait_ld();
                if (h == 3) {  // all TMEM reads of this accumulator are done
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(ACC_EMPTY(acc));
                }
                if (live) {
#pragma unroll
                    for (int c16 = 0; c16 < 4; ++c16) {
                        uint32_t o[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int col = c16 * 16 + 2 * i;
                            const __nv_bfloat162 r2 = *reinterpret_cast<const __nv_bfloat162*>(&res[h & 1][c16 * 8 + i]);
                            const float2 t = __bfloat1622float2(r2);
                            const float f0 = __uint_as_float(v[col]) + sbias[h * 64 + col] + t.x;
                            const float f1 = __uint_as_float(v[col + 1]) + sbias[h * 64 + col + 1] + t.y;
                            const __nv_bfloat162 ob = __floats2bfloat162_rn(fmaxf(f0, 0.f), fmaxf(f1, 0.f));
                            o[i] = *reinterpret_cast<const uint32_t*>(&ob);
                        }
                        stg256(out + off + h * 64 + c16 * 16, o);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(K::TMEM_COLS));
    }
}

template <typename Kern>
int launch_pdl(Kern kern, int threads, int grid, int smem, cudaStream_t s, const CUtensorMap& a_map, const CUtensorMap& w_map,
               const __nv_bfloat16* residual, __nv_bfloat16* out, const float* bias, int n_tiles, const HeadW& head,
               float* feat, const int* n_dev, int rev) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    RVS_CUDA(cudaLaunchKernelEx(&cfg, kern, a_map, w_map, residual, out, bias, n_tiles, head, feat, n_dev, rev));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

int launch_pdl_s(int grid, cudaStream_t s, const CUtensorMap& a_map, const CUtensorMap& w_map, const __nv_bfloat16* residual,
                 __nv_bfloat16* out, const float* bias, int n_tiles, const int* n_dev, int rev) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg2S::SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    RVS_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_tc2s_kernel, a_map, w_map, residual, out, bias, n_tiles, n_dev, rev));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

int encode_act_map(CUtensorMap* m, const void* ptr, int C /*channels of this buffer*/, int64_t n_tiles, int x_slots) {
    EncodeFn enc = get_encode();
    if (!enc) return fail(-9, "cuTensorMapEncodeTiled entry point not available");
    // tiled activation layout [tile][y][board][x][c]  (dims innermost first)
    const cuuint64_t dims[5] = {(cuuint64_t)C, 8, 2, 8, (cuuint64_t)n_tiles};
    const cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * 8, (cuuint64_t)C * 2 * 16, (cuuint64_t)C * 2 * 128};
    const cuuint32_t box[5] = {64, (cuuint32_t)x_slots, 2, 10, 1};  // 8: one box per dx (streamed kernel); 16: one box for all taps
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-9, "cuTensorMapEncodeTiled(activations) failed: %d", (int)r);
    return 0;
}

}  // namespace

int conv_tc_plan(ConvTcPlan& plan, const __nv_bfloat16* w, int C, int64_t max_batch, int cin) {
    plan.valid = false;
    plan.C = C;
    plan.max_batch = max_batch;
    if (cin <= 0) cin = C;
    if (!((C == 64 && cin == 64) || (C == 128 && (cin == 128 || cin == 64)) || (C == 256 && cin == 256)))
        return fail(-8, "tcgen05 convolution: unsupported shape %d -> %d", cin, C);
    EncodeFn enc = get_encode();
    if (!enc) return fail(-9, "cuTensorMapEncodeTiled entry point not available");
    Impl* im = plan.impl ? static_cast<Impl*>(plan.impl) : new Impl();
    plan.impl = im;
    im->n_act = 0;
    im->cin = cin;
    // weights [9*C rows (tap, cout)][cin] bf16, box = 64 cin x C/2 couts (the CTA's half of one tap)
    const cuuint64_t dims[2] = {(cuuint64_t)cin, (cuuint64_t)9 * C};
    const cuuint64_t strides[1] = {(cuuint64_t)cin * 2};
    const cuuint32_t es[2] = {1, 1};
    const cuuint32_t box2[2] = {64, (cuuint32_t)(C / 2)};
    CUresult r = enc(&im->w_map2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(w), dims, strides, box2, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-9, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
    // opt in to > 48 KB of dynamic shared memory on THIS device (function attributes are per device, and a
    // process may hold engines on several GPUs, so this is done per plan rather than once per process)
    if (C == 256) {
        RVS_CUDA(cudaFuncSetAttribute(conv3x3_tc2s_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2S::SMEM));
    } else if (C == 64) {
        RVS_CUDA(cudaFuncSetAttribute(conv3x3_tc2_kernel<64, 64, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<64, 64>::SMEM));
        RVS_CUDA(cudaFuncSetAttribute(conv3x3_tc2_kernel<64, 64, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<64, 64>::SMEM));
    } else if (cin == 64) {
        RVS_CUDA(cudaFuncSetAttribute(conv3x3_tc2_kernel<128, 64, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<128, 64>::SMEM));
        RVS_CUDA(cudaFuncSetAttribute(conv3x3_tc2_kernel<128, 64, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<128, 64>::SMEM));
    } else {
        RVS_CUDA(cudaFuncSetAttribute(conv3x3_tc2_kernel<128, 128, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<128, 128>::SMEM));
        RVS_CUDA(cudaFuncSetAttribute(conv3x3_tc2_kernel<128, 128, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<128, 128>::SMEM));
    }
    plan.valid = true;
    return 0;
}

int conv_tc_launch(const ConvTcPlan& plan, const __nv_bfloat16* in, const __nv_bfloat16* residual, __nv_bfloat16* out,
                   const float* bias, int64_t B, cudaStream_t s, const ConvHeadW* head, float* feat, const int* n_dev,
                   int max_ctas, int64_t cap_boards, int rev, int count_is_fresh) {
    if (!plan.valid || !plan.impl) return fail(-8, "tcgen05 convolution: no plan");
    Impl* im = static_cast<Impl*>(plan.impl);
    static const ConvHeadW zero_head = {};
    if (!(feat && head)) { feat = nullptr; head = &zero_head; }
    const int64_t cap = cap_boards > 0 ? cap_boards : plan.max_batch;  // boards addressable from `in`
    if (B > cap) return fail(-8, "tcgen05 convolution: batch %lld exceeds the buffer capacity %lld", (long long)B, (long long)cap);
    int slot = -1;
    for (int i = 0; i < im->n_act; ++i)
        if (im->act_ptr[i] == in && im->act_cap[i] == cap) slot = i;
    if (slot < 0) {
        if (im->n_act == Impl::kActSlots) im->n_act = 0;
        slot = im->n_act++;
        int rc = encode_act_map(&im->act_map[slot], in, im->cin, (cap + 1) / 2, plan.C == 256 ? 8 : kXSlots);  // tiles beyond the map are zero-filled by TMA
        if (rc) return rc;
        im->act_ptr[slot] = in;
        im->act_cap[slot] = cap;
    }
    const int n_tiles = (int)((B + 1) / 2);
    const int C = plan.C;
    // CTA pairs: grid = 2 x pairs, at most one CTA per SM (max_ctas < 148 leaves SMs to concurrent tree kernels)
    int pair_cap = (max_ctas > 0 && max_ctas < kNumSMs ? max_ctas : kNumSMs) / 2;
    if (pair_cap < 1) pair_cap = 1;
    int pairs = (n_tiles + 1) / 2;
    if (pairs > pair_cap) pairs = pair_cap;
    const CUtensorMap& am = im->act_map[slot];
    if (C == 256) return launch_pdl_s(2 * pairs, s, am, im->w_map2, residual, out, bias, n_tiles, n_dev, rev);
    if (C == 64) {
        if (feat) return launch_pdl(conv3x3_tc2_kernel<64, 64, true, false>, Cfg2<64, 64>::THREADS, 2 * pairs, Cfg2<64, 64>::SMEM, s, am, im->w_map2, residual, out, bias, n_tiles, *head, feat, n_dev, rev);
        return launch_pdl(conv3x3_tc2_kernel<64, 64, false, false>, Cfg2<64, 64>::THREADS, 2 * pairs, Cfg2<64, 64>::SMEM, s, am, im->w_map2, residual, out, bias, n_tiles, *head, feat, n_dev, rev);
    }
    if (im->cin == 64)  // first layer of a 128-filter tower: 64 (3 used) -> 128
    {
        if (count_is_fresh) return launch_pdl(conv3x3_tc2_kernel<128, 64, false, true>, Cfg2<128, 64>::THREADS, 2 * pairs, Cfg2<128, 64>::SMEM, s, am, im->w_map2, residual, out, bias, n_tiles, *head, nullptr, n_dev, rev);
        return launch_pdl(conv3x3_tc2_kernel<128, 64, false, false>, Cfg2<128, 64>::THREADS, 2 * pairs, Cfg2<128, 64>::SMEM, s, am, im->w_map2, residual, out, bias, n_tiles, *head, nullptr, n_dev, rev);
    }
    if (feat) return launch_pdl(conv3x3_tc2_kernel<128, 128, true, false>, Cfg2<128, 128>::THREADS, 2 * pairs, Cfg2<128, 128>::SMEM, s, am, im->w_map2, residual, out, bias, n_tiles, *head, feat, n_dev, rev);
    return launch_pdl(conv3x3_tc2_kernel<128, 128, false, false>, Cfg2<128, 128>::THREADS, 2 * pairs, Cfg2<128, 128>::SMEM, s, am, im->w_map2, residual, out, bias, n_tiles, *head, feat, n_dev, rev);
}

void conv_tc_destroy(ConvTcPlan& plan) {
    if (plan.impl) delete static_cast<Impl*>(plan.impl);
    plan.impl = nullptr;
    plan.valid = false;
}

#ifdef RVS_CONV_PROBE
extern "C" int rvs_debug_conv_probe(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, g_conv_probe, sizeof(long long) * 148 * 16);
}
#endif

bool conv_tc_can_fuse_head(const ConvTcPlan& plan) {
    return plan.valid && plan.impl && plan.C <= 128 && static_cast<Impl*>(plan.impl)->cin == plan.C;
}

}  // namespace rvs
