// ptx.cuh -- thin inline-PTX wrappers for the sm_100a features the kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM alloc / TMEM load / commit / fences).
// Every wrapper is one instruction (or one short sequence); no policy lives here.
#pragma once

#include <cuda.h> // CUtensorMap (type only; libcuda is resolved at run time, see tensormap.cuh)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nc
{

// Codes written to the per-handle error flag when a pipeline wait exceeds its budget.
enum : int
{
    KERR_NONE = 0,
    KERR_PRODUCER_EMPTY = 1,
    KERR_MMA_FULL = 2,
    KERR_MMA_TMEM_EMPTY = 3,
    KERR_EPI_TMEM_FULL = 4,
    KERR_SMEM_ALIGN = 5,
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Warp index as a value the compiler knows to be warp-uniform (a broadcast shuffle).  Everything a warp-specialised role derives
// from it -- ring slots, shared-memory descriptors, TMEM addresses -- can then live in uniform registers; computed from threadIdx
// alone those values are per-thread as far as the compiler can tell, and every tcgen05.mma / TMA instruction (which take uniform-
// register operands) is wrapped in an ELECT + R2UR.BROADCAST waterfall loop: ~160 cycles from one tcgen05.mma to the next in the INT8
// streaming kernels, where the MMAs are small.  Must be called by all threads of the warp, converged (kernel entry).
__device__ __forceinline__ int uniform_warp_idx()
{
#ifdef NETCUDA_NO_UNIFORM_WARP // A/B builds (NETCUDA_EXTRA_NVCC_FLAGS=-DNETCUDA_NO_UNIFORM_WARP NETCUDA_BUILD_TAG=... python build.py)
    return (int)(threadIdx.x >> 5);
#else
    return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
#endif
}

// ---- programmatic dependent launch ---------------------------------------------------------------
// Kernels of a forward pass are launched with cudaLaunchAttributeProgrammaticStreamSerialization: a kernel may start
// (set up barriers, allocate TMEM, prefetch descriptors) while its predecessor in the stream drains; griddep_wait()
// returns once the predecessor has completed and its writes are visible, and must precede every global access.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- mbarrier ------------------------------------------------------------------------------

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

// Make barrier initialisation visible to the async proxy (TMA, tcgen05.commit).
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// Order generic-proxy shared-memory writes before async-proxy reads (e.g. st.shared -> TMA store / UMMA).
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t"
                 "}"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return done != 0;
}

// Non-blocking probe of a phase (no hardware suspend): for threads that poll several barriers in turn.
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t"
                 "}"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return done != 0;
}

// Bounded wait.  A correct pipeline never gets near the budget (~4 s); a broken one records
// `code` in *err and traps, so a bug turns into a reported error instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int *err, int code)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity))
    {
        if ((++spins & 0x3FFu) == 0 && clock64() - t0 > 8000000000LL)
        {
            if (err) atomicExch(err, code);
            __threadfence_system();
            __trap();
        }
    }
}

// ---- TMA -----------------------------------------------------------------------------------

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// 2-D tiled load global -> shared; completion (bytes) is signalled on `bar`.
// c0 = coordinate along the contiguous (K) dimension, c1 = row coordinate, both in elements.
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst_smem), "l"(map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}

// Weights are read exactly once per launch: L2 evict-first, so that 128 MiB of them do not flush the biases (and whatever else the
// caller keeps in L2) on their way through.  It matters more than it looks: with ~29 MB of weight reads queued at the memory
// controllers, a bias load that misses L2 comes back 2-4 us later -- tools/stream_timeline.py showed the finalize of most layers
// waiting that long for 28 bias values.
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(policy)
                 : "memory");
}

// 2-D tiled store shared -> global (bulk async-group completion).  Out-of-bounds rows / columns of the
// box are clipped by the hardware, so ragged tile edges need no predicates.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src_smem, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src_smem), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, uint32_t src_smem, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// Same, but global += shared (element-wise add performed at L2, type taken from the tensor map).
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap *map, uint32_t src_smem, int c0, int c1)
{
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src_smem), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// All committed bulk stores of this thread have finished READING shared memory (the source may be reused).
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... all but the most recent one
__device__ __forceinline__ void tma_store_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Named barrier among `nthreads` threads of the CTA (id 0 is __syncthreads).
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- tcgen05: tensor memory ------------------------------------------------------------------

// Whole-warp (.sync.aligned).  Writes the TMEM base address of `ncols` columns to *dst_smem.
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish()
{
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// mbarrier arrive once every tcgen05.mma issued so far by this thread has completed.
__device__ __forceinline__ void tcgen05_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp receives row (lane) t.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr)
                 : "memory");
}
// Narrower forms of the same load: N consecutive 32-bit columns of the thread's row (N = 4, 8, 16).
template <int N>
__device__ __forceinline__ void tmem_ld_32xN(uint32_t taddr, uint32_t *r)
{
    static_assert(N == 4 || N == 8 || N == 16, "4, 8 or 16 columns");
    if constexpr (N == 4)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                     : "r"(taddr)
                     : "memory");
    else if constexpr (N == 8)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr)
                     : "memory");
    else
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr)
                     : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 16 consecutive 32-bit columns: thread t of the warp writes row (lane) t.
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t *r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                   "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- CTA pairs (cta_group::2): two SMs of one cluster share a 256-row UMMA ---------------------------

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// A shared::cta address is also the shared::cluster address of the same location in the executing CTA; clearing
// bit 24 turns it into the address of that location in the even (leader) CTA of the pair.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;

// TMA load issued by either CTA of a pair; the bytes are accounted on the LEADER CTA's mbarrier.
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst_smem, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst_smem), "l"(map), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1)
                 : "memory");
}
// Arrive on the leader CTA's copy of a barrier from either CTA of the pair.  Default (.release.cta) semantics:
// the TMEM reads this orders are fenced by tcgen05.fence::before_thread_sync; a .release.cluster arrive
// would cost a GPU-scope MEMBAR per tile and warp (measured: 38 % of the epilogue warps' time).
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_BIT_MASK) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair()
{
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// Arrive (once the pair's MMAs issued so far have completed) on the barrier at this smem offset in BOTH CTAs.
__device__ __forceinline__ void tcgen05_commit_pair(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}

// ---- distributed shared memory (any cluster; used by the split-K CTA pairs of mlp_umma_stream.cu) ---------------

// shared::cluster address of `smem_addr` (a shared::cta address of the executing CTA) in the CTA of rank `rank`.
__device__ __forceinline__ uint32_t mapa_shared(uint32_t smem_addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
// A store into another CTA's shared memory as an asynchronous operation that carries its own completion: 16 bytes land in the
// remote window and are counted (complete_tx) on the remote mbarrier -- the way TMA delivers data -- so the sender needs no fence and no arrive.
__device__ __forceinline__ void st_async_cluster_v4(uint32_t cluster_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t cluster_bar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(cluster_addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(cluster_bar)
                 : "memory");
}
// Bounded wait with cluster-scope acquire (data handed over through DSMEM).
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint32_t bar, uint32_t parity, int *err, int code)
{
    const long long t0 = clock64();
    uint32_t spins = 0;
    for (;;)
    {
        uint32_t done;
        asm volatile("{\n\t"
                     ".reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t"
                     "}"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (done) return;
        if ((++spins & 0x3FFu) == 0 && clock64() - t0 > 8000000000LL)
        {
            if (err) atomicExch(err, code);
            __threadfence_system();
            __trap();
        }
    }
}

// ---- tcgen05: descriptors and MMA ------------------------------------------------------------

// Shared-memory matrix descriptor for a K-major operand tile whose rows are 128 bytes and which
// was written by TMA with CU_TENSOR_MAP_SWIZZLE_128B (tile base 1024-byte aligned):
//   [0,14)  start address >> 4        [16,30) leading byte offset >> 4 (ignored for swizzled K-major; 1)
//   [32,46) stride byte offset >> 4 = 8 rows * 128 B = 1024 -> 64
//   [46,48) descriptor version = 1 (sm_100)            [61,64) layout type 2 = SWIZZLE_128B
// Advancing along K inside the 128-byte swizzle atom = adding (bytes >> 4) to the start address.
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Instruction descriptor (upper 32 bits of the "idesc" operand), dense, no negate, both operands K-major:
//   [4,6) D format (1 = F32, 2 = S32)   [7,10) A format   [10,13) B format
//   [17,23) N >> 3                      [24,29) M >> 4
// kind::f16: format 0 = F16, 1 = BF16;  kind::tf32: 2 = TF32;  kind::i8: 0 = U8, 1 = S8.
__host__ __device__ constexpr uint32_t umma_idesc(uint32_t d_fmt, uint32_t ab_fmt, uint32_t m, uint32_t n)
{
    return (d_fmt << 4) | (ab_fmt << 7) | (ab_fmt << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

constexpr uint32_t UMMA_IDESC_B_MN_MAJOR = 1u << 16; // B operand is N-contiguous ("MN-major") instead of K-contiguous

enum : int
{
    KIND_BF16 = 0,
    KIND_TF32 = 1,
    KIND_I8 = 2
};

// D[tmem] (+)= A[smem] * B[smem]^T.  Single-thread instruction; `accumulate` = 0 overwrites D.
template <int KIND>
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    if constexpr (KIND == KIND_BF16)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
                     : "memory");
    else if constexpr (KIND == KIND_TF32)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
                     : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
                     : "memory");
}

// CTA-pair form: issued by one thread of the leader CTA; M = 256 rows (128 per CTA), each CTA's smem holds
// its 128 rows of A and its half of the N rows of B at the same offsets.
template <int KIND>
__device__ __forceinline__ void umma_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    if constexpr (KIND == KIND_BF16)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
                     : "memory");
    else if constexpr (KIND == KIND_TF32)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
                     : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
                     : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (128 rows x 16 bf16 per instruction) is read from tensor
// memory, lane = row, two bf16 per 32-bit column.  Used for P.V in attention, where P never leaves TMEM.
__device__ __forceinline__ void umma_ts_bf16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
                 : "memory");
}

// ---- misc ------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); // .x = lo (low 16 bits), .y = hi
    return *reinterpret_cast<uint32_t *>(&v);
}

// Exact-erf GELU, x * Phi(x), written as  relu(x) - a * Phi(-a)  with  a = min(|x|, 6): no select between the two tails
// and no 1 - erf cancellation.  Phi(-a) = 2^P(a), P = degree-6 weighted least-squares fit of log2 Phi(-a) on [0, 6]
// (weight a * Phi(-a), i.e. the absolute error of the GELU itself); beyond 6 the term is < 6e-9.  Evaluated in fp32 the
// result is within 2.5e-7 (absolute) of x * (1 + erf(x / sqrt 2)) / 2 for all x -- the same as the classic erfc rational
// (Abramowitz-Stegun 7.1.26) -- with ONE MUFU operation and 11 issue slots per element instead of two and 14:
// the epilogue of the fc1 GEMM is bound by exactly these.
__device__ __forceinline__ float gelu_erf(float x)
{
    const float a = fminf(fabsf(x), 6.0f);
    float p = fmaf(3.3361295209033415e-05f, a, -0.0007681446732021868f);
    p = fmaf(p, a, 0.008066913112998009f);
    p = fmaf(p, a, -0.0533762164413929f);
    p = fmaf(p, a, -0.45880767703056335f);
    p = fmaf(p, a, -1.1511868238449097f);
    p = fmaf(p, a, -0.9999948740005493f);
    float h;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h) : "f"(p)); // Phi(-a)
    return fmaf(-a, h, fmaxf(x, 0.0f));
}

// Two elements at once on Blackwell's packed fp32 pipe (FFMA2 / FADD2: one issue slot for two IEEE-rn results, so the values are
// bit-identical to gelu_erf).  |x|, min and max have no packed form and stay scalar: 7.5 issue slots per element instead of 11.
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// (acc0 + bias0, acc1 + bias1) -> GELU of both
__device__ __forceinline__ void gelu_erf_x2(float acc0, float acc1, float bias0, float bias1, float &g0, float &g1)
{
    float x0, x1;
    unpack_f32x2(add_f32x2(pack_f32x2(acc0, acc1), pack_f32x2(bias0, bias1)), x0, x1);
    const float a0 = fminf(fabsf(x0), 6.0f), a1 = fminf(fabsf(x1), 6.0f);
    const uint64_t a = pack_f32x2(a0, a1);
    uint64_t p = fma_f32x2(pack_f32x2(3.3361295209033415e-05f, 3.3361295209033415e-05f), a,
                           pack_f32x2(-0.0007681446732021868f, -0.0007681446732021868f));
    p = fma_f32x2(p, a, pack_f32x2(0.008066913112998009f, 0.008066913112998009f));
    p = fma_f32x2(p, a, pack_f32x2(-0.0533762164413929f, -0.0533762164413929f));
    p = fma_f32x2(p, a, pack_f32x2(-0.45880767703056335f, -0.45880767703056335f));
    p = fma_f32x2(p, a, pack_f32x2(-1.1511868238449097f, -1.1511868238449097f));
    p = fma_f32x2(p, a, pack_f32x2(-0.9999948740005493f, -0.9999948740005493f));
    float p0, p1, h0, h1;
    unpack_f32x2(p, p0, p1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h0) : "f"(p0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h1) : "f"(p1));
    unpack_f32x2(fma_f32x2(pack_f32x2(-a0, -a1), pack_f32x2(h0, h1), pack_f32x2(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f))), g0, g1);
}

} // namespace nc
