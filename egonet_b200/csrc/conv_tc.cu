// tcgen05 implicit-GEMM convolution for sm_100a (fp16 in, fp32 accumulate in TMEM).
//
//   out[b,oh,ow,:] = act( sum_{r,s,c} in[b, oh*st+r-p, ow*st+s-p, c] * w[:, r,s,c] + bias (+ res) )
//
// GEMM view: M = output pixels (128 per CTA = one TMEM lane per pixel), N = output
// channels (one UMMA N tile of up to 256), K = taps x input channels.
//
// Data movement -- no im2col buffer, no halo staging, no padding branches:
//   * A (activations, NHWC fp16): for filter tap (r,s) the 128 x KC operand tile is
//     just the output tile's window shifted by (r-p, s-p).  One TMA tiled load of
//     the box {KC ch, TW, TH, TB} at signed coordinates fetches it; out-of-image
//     elements are zero-filled by the TMA unit, which IS the conv padding.
//     Stride-2 convs use a 5-D view [B][H/2][2][W/2][2*C] of the same buffer so
//     that the parity of the input row/column becomes a coordinate (still one
//     tiled load per tap).
//   * B (weights, [Cout][tap][Cin] fp16, BN folded): 2-D TMA box {KC, N}.
//   Both land in shared memory in the canonical K-major swizzled layout that the
//   UMMA shared-memory descriptors address, so the MMA issuer never touches data.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA
// issuer (one elected lane issues tcgen05.mma; tcgen05.commit recycles pipeline
// stages through mbarriers), warps 2-5 = epilogue (tcgen05.ld accumulator rows ->
// bias / residual / ReLU -> fp16 NHWC stores; optional fp32 NCHW heat-map copy).
// Several CTAs are resident per SM (small smem/TMEM footprints), so one CTA's
// epilogue overlaps its neighbours' main loops.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "common.h"
#include "kernels.h"

namespace egn {

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define EGN_TS(i) do { if (p.ts) p.ts[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 + (i)] = gtime(); } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// One lane of a converged warp.  The compiler recognises elect.sync and emits the guarded
// UTCHMMA / UTMALDG directly; predicating on (lane == 0) instead makes it wrap every such
// instruction in an ELECT + BRA.U.ANY waterfall loop (~10 dependent uniform-datapath instructions,
// ~150 cycles per MMA on the single issuing warp).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// Spin with a watchdog: a mis-programmed TMA/MMA pipeline traps (launch failure
// reported to the host) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4)
      : "memory");
}
// smem -> global tensor store (bulk async group); OOB parts of the box are clipped
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (TMA store) after the following barrier
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint4 lds_u4(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_u4(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// packed fp32 pair add (FADD2)
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long x = *reinterpret_cast<unsigned long long*>(&a), y = *reinterpret_cast<unsigned long long*>(&b), z;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(z) : "l"(x), "l"(y));
  return *reinterpret_cast<float2*>(&z);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// ---- CTA pair (cta_group::2) variants: two CTAs of a cluster drive one M = 256 MMA; each provides its own
// 128 rows of A and half of the N rows of B, at identical shared-memory offsets ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(const void* local, uint32_t rank) {
  uint32_t d;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(d) : "r"(smem_u32(local)), "r"(rank));
  return d;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads issued by either CTA of a pair; the transaction bytes are signalled on the barrier at `bar_cluster`
// (a shared::cluster address, normally the leader CTA's barrier)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// M = 256 MMA over the CTA pair, issued by ONE thread of the leader CTA
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this smem offset in BOTH CTAs once the pair's MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::f16, issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Two consecutive MMAs that read the SAME A tile (x_hi against w_hi, then against w_lo): the first keeps the tile
// in the tensor core's A collector (SASS .A_KEEP), the second takes it from there (.A_REUSE) instead of reading
// shared memory again -- an N = 192 MMA is close to the shared-memory read limit (A 4 KB + B 6 KB per 96 cycles).
__device__ __forceinline__ void umma_f16_keep_a(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_reuse_a(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t v[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t v[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// tcgen05.wait::ld that names the destination registers of the outstanding load as in/out operands, so
// the compiler cannot hoist arithmetic on them above the wait when the load was issued earlier.
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]),
                 "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]),
                 "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&v)[16]) { tmem_ld16(taddr, v); }
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&v)[32]) { tmem_ld32(taddr, v); }
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
// Asynchronous L2 prefetch of a contiguous global range (bytes: multiple of 16).
__device__ __forceinline__ void l2_prefetch_bulk(const void* gptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}

// K-major, swizzled shared-memory matrix descriptor (sm_100 "version 1"):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major, 1)
//   [32,46) stride byte offset >> 4 = 8 rows * swizzle span | [46,48) = 1 | [61,64) layout type
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t swizzle_bytes) {
  const uint64_t layout = swizzle_bytes == 128 ? 2 : (swizzle_bytes == 64 ? 4 : 6);
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * swizzle_bytes) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

// ---------------------------------------------------------------------------
// shared epilogue: TMEM accumulator rows -> bias (+residual) (+ReLU) -> fp16 NHWC
//
// One thread owns one accumulator row (= one output pixel).  The residual operand is the only
// dependent global read of the epilogue; with one 32-byte request per thread in flight it is pure
// latency (~0.7 us per 16 channels).  Work is therefore cut into groups of 32 channels and the
// residual of group g+1
// is requested (8 independent 16-byte loads per thread, double buffered in registers) before group g is
// converted and stored; the first group is requested before the
// accumulator barrier is even waited on, so it overlaps the whole main loop.
// ---------------------------------------------------------------------------
struct EpiArgs {
  const float* bias;
  const __half* res;
  __half* out;
  float* heatmap;
  const float* xs;
  const float* ys;
  int coord_maps, relu, Cout, Cout_p, OH, OW;
  int dbg;
  unsigned long long* ts;   // this CTA's stamp row or null
  // fp16x2 split storage (kernels.h): res / out are [pixel][hi: Cout_p | lo: Cout_p]; the tile's products
  // live in two accumulators `acc_stride` TMEM columns apart (hi*hi terms | the two cross terms) that are
  // summed here in round-to-nearest fp32
  int split;
  uint32_t acc_stride;
  uint32_t tile_stride;   // TMEM columns between consecutive M tiles (0 = n_tile)
};

// v -> (hi, lo) fp16 planes of 8 values each: hi = rn16(v), lo = rn16(v - hi)
__device__ __forceinline__ void split_store8(__half* p_hi, __half* p_lo, const float (&f)[8]) {
  uint4 h, l;
  __half2* h2 = reinterpret_cast<__half2*>(&h);
  __half2* l2 = reinterpret_cast<__half2*>(&l);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h2[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
    const float2 hf = __half22float2(h2[j]);
    l2[j] = __floats2half2_rn(f[2 * j] - hf.x, f[2 * j + 1] - hf.y);
  }
  *reinterpret_cast<uint4*>(p_hi) = h;
  *reinterpret_cast<uint4*>(p_lo) = l;
}
// 8 values of a split tensor: hi + lo
__device__ __forceinline__ void split_load8_add(const __half* p_hi, const __half* p_lo, float* f) {
  const uint4 h = __ldg(reinterpret_cast<const uint4*>(p_hi)), l = __ldg(reinterpret_cast<const uint4*>(p_lo));
  const __half2* h2 = reinterpret_cast<const __half2*>(&h);
  const __half2* l2 = reinterpret_cast<const __half2*>(&l);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 a = __half22float2(h2[j]), b = __half22float2(l2[j]);
    f[2 * j] += a.x + b.x;
    f[2 * j + 1] += a.y + b.y;
  }
}

struct EpiRow {      // where this thread's accumulator row of M-tile t lands
  bool valid;
  int b, oh, ow;
  size_t pix;
};

constexpr int kEpiGroup = 2;  // 16-column chunks per group

__device__ __forceinline__ void epi_prefetch(const EpiArgs& e, const EpiRow& r, int n, int nch, uint4 (&buf)[2 * kEpiGroup]) {
  if (!e.res || !r.valid || (e.dbg & 2) || e.split) return;     // split: the residual planes are read in epi_process
  const uint4* rp = reinterpret_cast<const uint4*>(e.res + r.pix * e.Cout_p + n);
#pragma unroll
  for (int i = 0; i < 2 * kEpiGroup; ++i)
    if (i < 2 * nch) buf[i] = __ldg(rp + i);
}

__device__ __forceinline__ void epi_process(const EpiArgs& e, const EpiRow& r, uint32_t taddr, int n, int nch,
                                            const uint4 (&buf)[2 * kEpiGroup]) {
  uint32_t v[kEpiGroup][16];
  if (!(e.dbg & 32)) {
#pragma unroll
    for (int c = 0; c < kEpiGroup; ++c)
      if (c < nch) tmem_ld16(taddr + 16u * c, v[c]);
    tmem_ld_wait();
  } else {
#pragma unroll
    for (int c = 0; c < kEpiGroup; ++c)
#pragma unroll
      for (int j = 0; j < 16; ++j) v[c][j] = 0;
  }
  if (e.split && e.acc_stride && !(e.dbg & 32)) {
    // second accumulator (cross terms) of the same columns
#pragma unroll
    for (int c = 0; c < kEpiGroup; ++c) {
      if (c >= nch) break;
      uint32_t x[16];
      tmem_ld16(taddr + e.acc_stride + 16u * c, x);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[c][j] = __float_as_uint(__uint_as_float(v[c][j]) + __uint_as_float(x[j]));
    }
  }
  if (!r.valid) return;
#pragma unroll
  for (int c = 0; c < kEpiGroup; ++c) {
    if (c >= nch) break;
    const int nn = n + 16 * c;
    float f[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 bq = reinterpret_cast<const float4*>(e.bias + nn)[j];
      f[4 * j + 0] = __uint_as_float(v[c][4 * j + 0]) + bq.x;
      f[4 * j + 1] = __uint_as_float(v[c][4 * j + 1]) + bq.y;
      f[4 * j + 2] = __uint_as_float(v[c][4 * j + 2]) + bq.z;
      f[4 * j + 3] = __uint_as_float(v[c][4 * j + 3]) + bq.w;
    }
    if (e.res && e.split) {
      if (!(e.dbg & 2)) {
        const __half* rp = e.res + r.pix * (size_t)(2 * e.Cout_p) + nn;
        split_load8_add(rp, rp + e.Cout_p, f);
        split_load8_add(rp + 8, rp + e.Cout_p + 8, f + 8);
      }
    } else if (e.res) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const __half2* r2 = reinterpret_cast<const __half2*>(&buf[2 * c + h]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 rf = __half22float2(r2[j]);
          f[8 * h + 2 * j] += rf.x;
          f[8 * h + 2 * j + 1] += rf.y;
        }
      }
    }
    if (e.relu) {
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    if (e.heatmap) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (nn + j < e.Cout) e.heatmap[(((size_t)r.b * e.Cout + nn + j) * e.OH + r.oh) * e.OW + r.ow] = f[j];
    }
    if (e.coord_maps) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (nn + j == e.Cout) f[j] = e.xs[r.ow];
        if (nn + j == e.Cout + 1) f[j] = e.ys[r.oh];
      }
    }
    if (e.split) {
      if (!(e.dbg & 1)) {
        __half* op = e.out + r.pix * (size_t)(2 * e.Cout_p) + nn;
        const float (&f0)[8] = *reinterpret_cast<const float (*)[8]>(&f[0]);
        const float (&f1)[8] = *reinterpret_cast<const float (*)[8]>(&f[8]);
        split_store8(op, op + e.Cout_p, f0);
        split_store8(op + 8, op + e.Cout_p + 8, f1);
      }
      continue;
    }
    uint4 o[2];
    __half2* o2 = reinterpret_cast<__half2*>(o);
#pragma unroll
    for (int j = 0; j < 8; ++j) o2[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
    uint4* op = reinterpret_cast<uint4*>(e.out + r.pix * e.Cout_p + nn);
    if (!(e.dbg & 1)) {
      op[0] = o[0];
      op[1] = o[1];
    }
  }
}

// Runs the grouped, double-buffered epilogue over T accumulator tiles of n_tile columns.
// row_of(t) maps this thread's TMEM lane of tile t to its output pixel.  Group bookkeeping is
// incremental (no integer division on the critical path).
struct EpiCursor {
  int t, gi;       // tile, group inside the tile
  EpiRow row;
};

template <typename RowFn>
__device__ __forceinline__ void epi_run(const EpiArgs& e, uint32_t tmem_lane_base, int T, int n_tile, int n0,
                                        uint64_t* tmem_full_bar, RowFn row_of, uint32_t parity = 0, int half = 0,
                                        int halves = 1) {
  // With halves == 2 two warps share each TMEM lane quarter and split the (tile, group) items of a
  // window in a checkerboard ((t + group) & 1), which balances 32/16-column groups across the pair.
  const int nch_total = n_tile >> 4;
  const int groups = (nch_total + kEpiGroup - 1) / kEpiGroup;
  uint4 bufA[2 * kEpiGroup], bufB[2 * kEpiGroup];
  auto step = [&](EpiCursor& c) {             // next group; recompute the row only when the tile changes
    if (++c.gi == groups) {
      c.gi = 0;
      ++c.t;
      if (c.t < T) c.row = row_of(c.t);
    }
  };
  auto mine = [&](const EpiCursor& c) { return halves == 1 || ((c.t + c.gi) & 1) == half; };
  auto advance = [&](EpiCursor& c) {
    step(c);
    while (c.t < T && !mine(c)) step(c);      // at most two foreign items in a row (tile boundary)
  };
  auto nch_of = [&](const EpiCursor& c) { return min(kEpiGroup, nch_total - c.gi * kEpiGroup); };
  auto n_of = [&](const EpiCursor& c) { return n0 + 16 * kEpiGroup * c.gi; };
  const int tstride = e.tile_stride ? (int)e.tile_stride : n_tile;
  auto ta_of = [&](const EpiCursor& c) { return tmem_lane_base + (uint32_t)(c.t * tstride + 16 * kEpiGroup * c.gi); };
  EpiCursor ca{0, 0, row_of(0)};
  while (ca.t < T && !mine(ca)) step(ca);
  if (ca.t < T) epi_prefetch(e, ca.row, n_of(ca), nch_of(ca), bufA);
  mbar_wait(tmem_full_bar, parity);
  tc_fence_after();
  if (e.ts && (threadIdx.x & 127) == 64) e.ts[4] = gtime();
  while (ca.t < T) {
    EpiCursor cb = ca;
    advance(cb);
    const bool has_b = cb.t < T;
    if (has_b) epi_prefetch(e, cb.row, n_of(cb), nch_of(cb), bufB);
    epi_process(e, ca.row, ta_of(ca), n_of(ca), nch_of(ca), bufA);
    if (!has_b) break;
    ca = cb;
    advance(ca);
    if (ca.t < T) epi_prefetch(e, ca.row, n_of(ca), nch_of(ca), bufA);
    epi_process(e, cb.row, ta_of(cb), n_of(cb), nch_of(cb), bufB);
  }
}

// ---------------------------------------------------------------------------
// compact epilogue of the persistent kernel (plain layers: bias [+ residual] [+ ReLU] -> fp16 NHWC)
//
// The generic epilogue above is ~3000 straight-line instructions per window (two register buffers x two
// column chunks x the head extras); a warp walking through it once per window runs at ~0.25 IPC.  This
// one is a single-item loop body without the head code.  An item is 16 accumulator columns of one M
// tile; kParts (4) warps share a TMEM lane quarter and take items idx = part, part + kParts, ...
// While item i is converted, the TMEM load of item i+1 and the residual loads of items i+1 and i+2
// (2 x 32 bytes per thread) are in flight.
// ---------------------------------------------------------------------------
struct EpiLite {
  const __half* res;
  __half* out;
  uint32_t bias_s;     // shared-space address of this CTA's bias slice (n_tile floats)
  int relu, Cout_p, dbg;
  int split;           // fp16x2 storage: res / out are [pixel][hi: Cout_p | lo: Cout_p]
};

struct LiteItem {
  int t, g;            // M tile, 16-column chunk inside the tile
  bool valid;
  size_t pix;
};

__device__ __forceinline__ void lite_res_load(const EpiLite& e, const LiteItem& it, int n, uint4 (&rb)[2]) {
  if (!e.res || !it.valid || (e.dbg & 2) || e.split) return;     // split: read in lite_store
  const uint4* rp = reinterpret_cast<const uint4*>(e.res + it.pix * e.Cout_p + n);
  rb[0] = __ldg(rp);
  rb[1] = __ldg(rp + 1);
}

__device__ __forceinline__ void lite_store(const EpiLite& e, const LiteItem& it, int n, int nl, const uint32_t (&v)[16],
                                           const uint4 (&rb)[2]) {
  if (!it.valid) return;
  float f[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 bq = lds_f4(e.bias_s + (uint32_t)(nl + 4 * j) * 4u);
    f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + bq.x;
    f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + bq.y;
    f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + bq.z;
    f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + bq.w;
  }
  if (e.res && e.split) {
    if (!(e.dbg & 2)) {
      const __half* rp = e.res + it.pix * (size_t)(2 * e.Cout_p) + n;
      split_load8_add(rp, rp + e.Cout_p, f);
      split_load8_add(rp + 8, rp + e.Cout_p + 8, f + 8);
    }
  } else if (e.res) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const __half2* r2 = reinterpret_cast<const __half2*>(&rb[h]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 rf = __half22float2(r2[j]);
        f[8 * h + 2 * j] += rf.x;
        f[8 * h + 2 * j + 1] += rf.y;
      }
    }
  }
  if (e.relu) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
  }
  if (e.split) {
    if (!(e.dbg & 1)) {
      __half* op = e.out + it.pix * (size_t)(2 * e.Cout_p) + n;
      const float (&f0)[8] = *reinterpret_cast<const float (*)[8]>(&f[0]);
      const float (&f1)[8] = *reinterpret_cast<const float (*)[8]>(&f[8]);
      split_store8(op, op + e.Cout_p, f0);
      split_store8(op + 8, op + e.Cout_p + 8, f1);
    }
    return;
  }
  uint4 o[2];
  __half2* o2 = reinterpret_cast<__half2*>(o);
#pragma unroll
  for (int j = 0; j < 8; ++j) o2[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
  if (!(e.dbg & 1)) {
    uint4* op = reinterpret_cast<uint4*>(e.out + it.pix * e.Cout_p + n);
    op[0] = o[0];
    op[1] = o[1];
  }
}

// K-split accumulators: the K loop of a tile is dealt round-robin over `ksplit` TMEM accumulators (consecutive
// tcgen05.mma then never wait for each other's result: 62 -> 44 cycles per 128x48x16 MMA between 1 and 8
// accumulators, profiles/r01k_umma_rate.log); the epilogue adds the partial sums.  v holds accumulator 0 of the
// 16 columns at `taddr`; the others sit `acc_stride` columns apart.
__device__ __forceinline__ void add_split_accumulators(uint32_t (&v)[16], uint32_t taddr, int ksplit, uint32_t acc_stride) {
  for (int a = 1; a < ksplit; ++a) {
    uint32_t x[16];
    tmem_ld(taddr + (uint32_t)a * acc_stride, x);
    tmem_ld_wait_dep(x);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(x[i]));
  }
}

template <int kParts, typename RowFn>
__device__ __forceinline__ void epi_window_lite(const EpiLite& e, uint32_t tmem_lane_base, int T, int n_tile, int n0,
                                                uint64_t* full_bar, uint32_t parity, int half, RowFn row_of,
                                                unsigned long long* ts, int ksplit = 1, int stacked = 0) {
  // stacked (fp16x2, persistent kernel): tile t keeps [H | L] in adjacent column ranges of n_tile each
  const int per_tile = n_tile >> 4;
  const uint32_t acc_stride = (uint32_t)(stacked ? n_tile : T * n_tile);
  const int tstride = stacked ? 2 * n_tile : n_tile;
  auto next = [&](const LiteItem& it) {              // item + kParts; refresh the row when the tile changes
    LiteItem n = it;
    n.g += kParts;
    while (n.g >= per_tile) {
      n.g -= per_tile;
      ++n.t;
    }
    if (n.t != it.t && n.t < T) {
      const EpiRow r = row_of(n.t);
      n.valid = r.valid;
      n.pix = r.pix;
    }
    return n;
  };
  auto taddr = [&](const LiteItem& it) { return tmem_lane_base + (uint32_t)(it.t * tstride + 16 * it.g); };

  LiteItem cur{-1, half - kParts + per_tile, false, 0};   // one step before the first item of tile 0
  cur = next(cur);
  LiteItem n1 = next(cur), n2 = next(n1);
  uint4 r0[2] = {}, r1[2] = {}, r2[2] = {};
  uint32_t v0[16], v1[16];
  if (cur.t < T) lite_res_load(e, cur, n0 + 16 * cur.g, r0);
  if (n1.t < T) lite_res_load(e, n1, n0 + 16 * n1.g, r1);
  mbar_wait(full_bar, parity);
  tc_fence_after();
  if (ts && (threadIdx.x & 127) == 64) ts[4] = gtime();
  if (cur.t < T) tmem_ld(taddr(cur), v1);
#pragma unroll 1
  while (cur.t < T) {
    if (n2.t < T) lite_res_load(e, n2, n0 + 16 * n2.g, r2);
    tmem_ld_wait_dep(v1);
#pragma unroll
    for (int i = 0; i < 16; ++i) v0[i] = v1[i];
    if (n1.t < T) tmem_ld(taddr(n1), v1);
    if (ksplit > 1) add_split_accumulators(v0, taddr(cur), ksplit, acc_stride);
    lite_store(e, cur, n0 + 16 * cur.g, 16 * cur.g, v0, r0);
    r0[0] = r1[0]; r0[1] = r1[1];
    r1[0] = r2[0]; r1[1] = r2[1];
    cur = n1;
    n1 = n2;
    n2 = next(n2);
  }
}

// ---------------------------------------------------------------------------
// staged epilogue of the persistent kernel
//
// Direct 16-byte-per-lane global accesses at a row pitch of Cout_p * 2 bytes touch one L1 line per lane
// (~6x the wavefronts of a coalesced access) and those wavefronts compete with the tensor core's operand
// reads for the same shared-memory / L1 data path: with the stores enabled the MMA phase of a
// 48-channel window stretches from 3750 to 6000 SM cycles (profiles/r01_v3b_epilogue_ablation.log).
// Here the residual window arrives by TMA in a staging buffer laid out [block][pixel][cb channels], every
// thread converts its 16-column items IN PLACE (read residual, write output at the same address) and
// the buffer leaves by one TMA tensor store per block, which also clips rows beyond the image / batch.
// ---------------------------------------------------------------------------
struct EpiStage {
  uint32_t base;       // shared-space address of this window's staging buffer
  uint32_t bias_s;
  uint32_t blk_bytes;  // bytes of one channel block (rows_stage * cb * 2, 1024-aligned)
  int cb;              // channels per block: 64 (128-byte rows, SWIZZLE_128B) or 48 (96-byte rows, no swizzle)
  int has_res, relu, dbg;
  uint32_t lo_off;     // fp16x2 storage: byte offset from a hi-plane block to its lo-plane block (0 = plain fp16)
};

struct StageRow {
  bool valid;          // run position is an output pixel of this window
  uint32_t srow;       // its row in the staging buffer
};

template <int kParts, typename RowFn>
__device__ __forceinline__ void epi_window_staged(const EpiStage& e, uint32_t tmem_lane_base, int T, int n_tile,
                                                  int part, RowFn row_of, int ksplit = 1, int stacked = 0) {
  const int per_tile = n_tile >> 4;
  const uint32_t acc_stride = (uint32_t)(stacked ? n_tile : T * n_tile);
  const int tstride = stacked ? 2 * n_tile : n_tile;
  const uint32_t pitch = (uint32_t)e.cb * 2u;
  int t = -1, g = part - kParts + per_tile, t_row = -1;
  StageRow r{false, 0};
  auto step = [&]() {
    g += kParts;
    while (g >= per_tile) {
      g -= per_tile;
      ++t;
    }
  };
  step();
  // one item: accumulators (already loaded) + bias [+ residual from the staging buffer] -> fp16 in place
  auto item = [&](const uint32_t (&v)[16], const StageRow& sr, int cg) {
    if (!sr.valid || (e.dbg & 256)) return;
    const int col = 16 * cg;
    const int blk = e.cb == 64 ? (col >> 6) : col / e.cb;
    const uint32_t ch = (uint32_t)(col - blk * e.cb) >> 3;                    // first 16-byte chunk (even)
    const uint32_t rowb = e.base + (uint32_t)blk * e.blk_bytes + sr.srow * pitch;
    const uint32_t sw = e.cb == 64 ? (sr.srow & 7u) : 0u;
    const uint32_t a0 = rowb + ((ch ^ sw) << 4), a1 = rowb + (((ch + 1u) ^ sw) << 4);
    float2 f[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 bq = lds_f4(e.bias_s + (uint32_t)(col + 4 * j) * 4u);
      f[2 * j] = add2(make_float2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), make_float2(bq.x, bq.y));
      f[2 * j + 1] =
          add2(make_float2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), make_float2(bq.z, bq.w));
    }
    if (e.has_res) {
      uint4 q[2];
      q[0] = lds_u4(a0);
      q[1] = lds_u4(a1);
      const __half2* r2 = reinterpret_cast<const __half2*>(q);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = add2(f[j], __half22float2(r2[j]));
      if (e.lo_off) {
        q[0] = lds_u4(a0 + e.lo_off);
        q[1] = lds_u4(a1 + e.lo_off);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = add2(f[j], __half22float2(r2[j]));
      }
    }
    if (e.lo_off) {
      // split output: relu in fp32, hi = rn16(v), lo = rn16(v - hi), both planes converted in place
      uint4 oh[2], ol[2];
      __half2* h2 = reinterpret_cast<__half2*>(oh);
      __half2* l2 = reinterpret_cast<__half2*>(ol);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float2 x = f[j];
        if (e.relu) {
          x.x = fmaxf(x.x, 0.f);
          x.y = fmaxf(x.y, 0.f);
        }
        h2[j] = __float22half2_rn(x);
        const float2 hf = __half22float2(h2[j]);
        l2[j] = __floats2half2_rn(x.x - hf.x, x.y - hf.y);
      }
      sts_u4(a0, oh[0]);
      sts_u4(a1, oh[1]);
      sts_u4(a0 + e.lo_off, ol[0]);
      sts_u4(a1 + e.lo_off, ol[1]);
      return;
    }
    uint4 o[2];
    __half2* o2 = reinterpret_cast<__half2*>(o);
#pragma unroll
    for (int j = 0; j < 8; ++j) o2[j] = __float22half2_rn(f[j]);
    if (e.relu) {
      const __half2 z = __float2half2_rn(0.f);
#pragma unroll
      for (int j = 0; j < 8; ++j) o2[j] = __hmax2(o2[j], z);     // relu(round(x)) == round(relu(x))
    }
    sts_u4(a0, o[0]);
    sts_u4(a1, o[1]);
  };
  auto row_for = [&](int tt) {
    if (tt != t_row) {
      r = row_of(tt);
      t_row = tt;
    }
  };
  // software pipeline over this warp's items with two register buffers: the TMEM load of the next item
  // is in flight while the current one is converted
  uint32_t vA[16], vB[16];
  if (t < T) tmem_ld(tmem_lane_base + (uint32_t)(t * tstride + 16 * g), vA);
#pragma unroll 1
  while (t < T) {
    int ct = t, cg = g;
    step();
    tmem_ld_wait_dep(vA);
    if (t < T) tmem_ld(tmem_lane_base + (uint32_t)(t * tstride + 16 * g), vB);
    if (ksplit > 1) add_split_accumulators(vA, tmem_lane_base + (uint32_t)(ct * tstride + 16 * cg), ksplit, acc_stride);
    row_for(ct);
    item(vA, r, cg);
    if (t >= T) break;
    ct = t;
    cg = g;
    step();
    tmem_ld_wait_dep(vB);
    if (t < T) tmem_ld(tmem_lane_base + (uint32_t)(t * tstride + 16 * g), vA);
    if (ksplit > 1) add_split_accumulators(vB, tmem_lane_base + (uint32_t)(ct * tstride + 16 * cg), ksplit, acc_stride);
    row_for(ct);
    item(vB, r, cg);
  }
}

// ---------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------
constexpr int kTcThreads = 192;
constexpr int kMaxStages = 8;

struct TcParams {
  // problem
  int B, OH, OW, Cout_p, Cout, Cin_p;
  int taps, ksize, stride, pad, relu;
  // tiling
  int TW, TH, TB;            // output tile = TB x TH x TW pixels (<= 128)
  int tiles_w, tiles_h;      // tiles per image along w / h
  int n_tile;                // UMMA N (output channels per CTA)
  int kc;                    // channels per pipeline stage (= swizzle bytes / 2)
  int kchunks;               // ceil(Cin_p / kc)
  int cin_k;                 // kchunks * kc: per-tap K extent of the zero-padded weight matrix
  int stages;
  uint32_t a_bytes, b_bytes; // TMA transaction bytes per stage
  uint32_t tmem_cols;
  // epilogue operands
  const float* bias;
  const __half* res;
  __half* out;
  float* heatmap;
  const float* xs;
  const float* ys;
  int coord_maps;
  // fp16x2 split storage: the activation tensor has cin_a = 2 * Cin_p physical channels ([hi | lo] planes), nh =
  // Cin_p / 16 K16 slices per plane.  A pipeline stage is "fat": for one (tap, kc-channel chunk of the LOGICAL input
  // channels) it holds all four operands -- the x_hi box, the x_lo box and the row-stacked weight tile
  // [w_hi (n_tile rows) | w_lo (n_tile rows)] (the dense stacked weight matrix of PersistParams) -- so every operand
  // byte is fetched once per tap (the previous [x_hi | x_lo] x [w_hi | w_hi], then x_hi x w_lo staging fetched x_hi and
  // w_hi twice: the per-tap kernel is bound by the bytes its shared-memory ring can keep in flight, so 1.5x the MMA
  // work per byte is 1.5x the speed).  Per K16 slice: N = 2 n_tile <= 256: one MMA x_hi x [w_hi | w_lo] -> [H | L] and
  // one x_lo x w_hi -> L; wider tiles: x_hi x w_hi -> H, x_hi x w_lo -> L, x_lo x w_hi -> L.  H (hi*hi) lives in
  // TMEM columns [0, n_tile), L (cross terms) in [n_tile, 2 n_tile); the epilogue adds them.
  int split, cin_a, nh;
  int keep_a;    // 3-MMA form: the second x_hi MMA takes the A tile from the collector (umma_f16_keep_a)
  // staged TMA epilogue over the dead stage ring (TapWinParams::staged): staging box = the output tile TW x TH x TB
  int staged, cb, nblk_plane;
  uint32_t blk_bytes;
  int grouped;   // fp16x2: issue the MMAs of a stage grouped by accumulator
};

template <int SW>
__global__ void __launch_bounds__(kTcThreads, 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_res, const __grid_constant__ CUtensorMap map_out,
               const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte aligned operand ring (swizzle atoms are address based)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_plane = 128u * SW;                                  // one activation box
  const uint32_t a_stage = p.split ? 2u * a_plane : a_plane;           // split: [x_hi box | x_lo box]
  const uint32_t b_stage = ((uint32_t)p.n_tile * SW * (p.split ? 2u : 1u) + 1023u) & ~1023u;   // split: [w_hi | w_lo] rows
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + (size_t)p.stages * a_stage;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + (size_t)p.stages * b_stage);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kMaxStages;
  uint64_t* res_full = tmem_full_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 1);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 2);       // [n_tile]; offset 18 * 8 + 8 = 152 ... +8 below: 16-byte aligned
  s_bias += 2;                                                   // 160

  // warp index broadcast through a shuffle so the compiler can prove the role branches warp-uniform
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  // tile coordinates
  int tile = blockIdx.x;
  const int tw_i = tile % p.tiles_w;
  tile /= p.tiles_w;
  const int th_i = tile % p.tiles_h;
  const int tb_i = tile / p.tiles_h;
  const int ow0 = tw_i * p.TW, oh0 = th_i * p.TH, b0 = tb_i * p.TB;
  const int n0 = blockIdx.y * p.n_tile;
  if (p.staged)
    for (int i = threadIdx.x; i < p.n_tile; i += (int)blockDim.x) s_bias[i] = __ldg(p.bias + n0 + i);

  if (threadIdx.x == 0) {
    pdl_trigger();
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (p.staged) {
      tma_prefetch_desc(&map_res);
      tma_prefetch_desc(&map_out);
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(res_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();   // everything above is independent of the previous kernel (common.h, PDL rules)

  // warp-uniform role loops, instructions predicated to lane 0 (see conv_run_kernel)
  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t stage = 0, phase = 0;
    int kcoord = 0;                                  // B coordinate of the stage: stages are stored back to back
    int tap = 0;
    for (int r = 0; r < p.ksize; ++r) {
      for (int q = 0; q < p.ksize; ++q, ++tap) {
        // stride 2: input row = 2*(oh + dh) + hp, input col = 2*(ow + dw) + wp
        const int er = r - p.pad, eq = q - p.pad;
        const int hp = er & 1, dh = (er - hp) / 2;
        const int wp = eq & 1, dw = (eq - wp) / 2;
        for (int kcx = 0; kcx < p.kchunks; ++kcx, kcoord += p.kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          const int c0 = kcx * p.kc;
          if (elect_one()) {
            uint8_t* sa = smem_a + stage * a_stage;
            uint8_t* sb = smem_b + stage * b_stage;
            mbar_expect_tx(&full_bar[stage], (p.split ? 2u : 1u) * (p.a_bytes + p.b_bytes));
            if (p.stride == 1)
              tma_load_4d(sa, &map_a, &full_bar[stage], c0, ow0 + q - p.pad, oh0 + r - p.pad, b0);
            else
              tma_load_5d(sa, &map_a, &full_bar[stage], wp * p.cin_a + c0, ow0 + dw, hp, oh0 + dh, b0);
            if (p.split) {
              // the lo-plane box of the same channels, and the stacked weight tile of (tap, chunk): dense K index
              if (p.stride == 1)
                tma_load_4d(sa + a_plane, &map_a, &full_bar[stage], p.Cin_p + c0, ow0 + q - p.pad, oh0 + r - p.pad, b0);
              else
                tma_load_5d(sa + a_plane, &map_a, &full_bar[stage], wp * p.cin_a + p.Cin_p + c0, ow0 + dw, hp, oh0 + dh, b0);
              tma_load_2d(sb, &map_b, &full_bar[stage], tap * p.Cin_p + c0, n0);
              tma_load_2d(sb + p.b_bytes, &map_b, &full_bar[stage], tap * p.Cin_p + c0, p.Cout_p + n0);
            } else {
              tma_load_2d(sb, &map_b, &full_bar[stage], kcoord, n0);
            }
          }
          if (++stage == (uint32_t)p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    if (p.staged && p.res) {
      // every MMA has completed: the stage ring is dead, the residual tile lands on top of it
      mbar_wait(tmem_full_bar, 0);
      if (elect_one()) {
        const int nblk = (p.split ? 2 : 1) * p.nblk_plane;
        mbar_expect_tx(res_full, (uint32_t)nblk * (uint32_t)(p.TW * p.TH * p.TB * p.cb * 2));
        for (int k = 0; k < nblk; ++k) {
          const int ch = k < p.nblk_plane ? n0 + k * p.cb : p.Cout_p + n0 + (k - p.nblk_plane) * p.cb;
          tma_load_4d(smem + (size_t)k * p.blk_bytes, &map_res, res_full, ch, ow0, oh0, b0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: D=f32, A=B=f16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
    const uint32_t idesc = (1u << 4) | ((uint32_t)(p.n_tile >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc_w = (1u << 4) | ((uint32_t)(p.n_tile >> 2) << 17) | ((128u >> 4) << 24);   // N = 2 n_tile
    const bool wide = p.split && 2 * p.n_tile <= 256;
    const uint64_t desc_hi = make_smem_desc(0, SW);
    const uint32_t a_addr0 = smem_u32(smem_a), b_addr0 = smem_u32(smem_b);
    const int n_iters = p.taps * p.kchunks;
    uint32_t stage = 0, phase = 0;
    int kcx = 0;
    for (int it = 0; it < n_iters; ++it) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint64_t adesc = desc_hi | (uint64_t)(((a_addr0 + stage * a_stage) & 0x3FFFFu) >> 4);
      const uint64_t bdesc = desc_hi | (uint64_t)(((b_addr0 + stage * b_stage) & 0x3FFFFu) >> 4);
      if (elect_one()) {
        if (!p.split) {
#pragma unroll
          for (int k = 0; k < SW / 32; ++k)      // +32 bytes (16 fp16) along K inside the swizzle span: start-address field += 2
            umma_f16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it | k) ? 1u : 0u);
        } else {
          const uint64_t adesc_lo = adesc + (uint64_t)(a_plane >> 4);
          const uint64_t bdesc_lo = bdesc + (uint64_t)(p.b_bytes >> 4);
          if (p.grouped) {
            // same-accumulator runs (see conv_tapwin_kernel)
#pragma unroll
            for (int k = 0; k < SW / 32; ++k) {
              if (kcx * (SW / 32) + k >= p.nh) continue;
              umma_f16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), wide ? idesc_w : idesc, (it | k) ? 1u : 0u);
            }
            if (!wide) {
#pragma unroll
              for (int k = 0; k < SW / 32; ++k) {
                if (kcx * (SW / 32) + k >= p.nh) continue;
                umma_f16(tmem_base + (uint32_t)p.n_tile, adesc + (uint64_t)(2 * k), bdesc_lo + (uint64_t)(2 * k), idesc, (it | k) ? 1u : 0u);
              }
            }
#pragma unroll
            for (int k = 0; k < SW / 32; ++k) {
              if (kcx * (SW / 32) + k >= p.nh) continue;
              umma_f16(tmem_base + (uint32_t)p.n_tile, adesc_lo + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
            }
          } else
#pragma unroll
          for (int k = 0; k < SW / 32; ++k) {
            if (kcx * (SW / 32) + k >= p.nh) continue;             // padding slice of the last chunk
            const uint32_t acc = (it | k) ? 1u : 0u;               // the first slice initialises H and L
            const uint64_t ko = (uint64_t)(2 * k);
            if (wide) {
              umma_f16(tmem_base, adesc + ko, bdesc + ko, idesc_w, acc);                               // [H | L] += x_hi [w_hi | w_lo]
              umma_f16(tmem_base + (uint32_t)p.n_tile, adesc_lo + ko, bdesc + ko, idesc, 1u);          // L += x_lo w_hi
            } else {
              if (p.keep_a) {
                umma_f16_keep_a(tmem_base, adesc + ko, bdesc + ko, idesc, acc);                        // H += x_hi w_hi
                umma_f16_reuse_a(tmem_base + (uint32_t)p.n_tile, adesc + ko, bdesc_lo + ko, idesc, acc);   // L += x_hi w_lo
              } else {
                umma_f16(tmem_base, adesc + ko, bdesc + ko, idesc, acc);                               // H += x_hi w_hi
                umma_f16(tmem_base + (uint32_t)p.n_tile, adesc + ko, bdesc_lo + ko, idesc, acc);       // L += x_hi w_lo
              }
              umma_f16(tmem_base + (uint32_t)p.n_tile, adesc_lo + ko, bdesc + ko, idesc, 1u);          // L += x_lo w_hi
            }
          }
        }
        umma_commit(&empty_bar[stage]);  // frees this stage once the MMAs above have read it
      }
      if (++kcx == p.kchunks) kcx = 0;
      if (++stage == (uint32_t)p.stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
    if (elect_one()) umma_commit(tmem_full_bar);    // accumulator complete
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;     // accumulator row = pixel index inside the tile
    const int tw = row % p.TW;
    const int th = (row / p.TW) % p.TH;
    const int tb = row / (p.TW * p.TH);
    EpiRow er;
    er.b = b0 + tb;
    er.oh = oh0 + th;
    er.ow = ow0 + tw;
    er.valid = row < p.TW * p.TH * p.TB && er.b < p.B && er.oh < p.OH && er.ow < p.OW;
    er.pix = ((size_t)er.b * p.OH + er.oh) * p.OW + er.ow;
    if (p.staged) {
      const int nblk = (p.split ? 2 : 1) * p.nblk_plane;
      const EpiStage es{smem_u32(smem), smem_u32(s_bias), p.blk_bytes, p.cb, p.res != nullptr, p.relu, 0,
                        p.split ? (uint32_t)p.nblk_plane * p.blk_bytes : 0u};
      mbar_wait(tmem_full_bar, 0);
      if (p.res) mbar_wait(res_full, 0);
      tc_fence_after();
      epi_window_staged<1>(es, tmem_base + ((uint32_t)(quarter * 32) << 16), 1, p.n_tile, 0, [&](int) {
        StageRow sr;
        sr.valid = row < p.TW * p.TH * p.TB;       // pixels past the image / batch are clipped by the TMA store
        sr.srow = (uint32_t)row;                   // the staging box is the tile, row-major [TB][TH][TW]
        return sr;
      }, p.split ? 2 : 1, p.split);
      fence_proxy_async();
      asm volatile("bar.sync 1, 128;" ::: "memory");        // the four epilogue warps
      if (warp == 2 && elect_one()) {
        for (int k = 0; k < nblk; ++k) {
          const int ch = k < p.nblk_plane ? n0 + k * p.cb : p.Cout_p + n0 + (k - p.nblk_plane) * p.cb;
          tma_store_4d(&map_out, smem + (size_t)k * p.blk_bytes, ch, ow0, oh0, b0);
        }
        bulk_commit();
        bulk_wait0();
      }
    } else {
      EpiArgs e{p.bias, p.res, p.out, p.heatmap, p.xs, p.ys, p.coord_maps, p.relu, p.Cout, p.Cout_p, p.OH, p.OW, 0, nullptr,
                p.split, (uint32_t)p.n_tile, 0u};
      epi_run(e, tmem_base + ((uint32_t)(quarter * 32) << 16), 1, p.n_tile, n0, tmem_full_bar,
              [&](int) { return er; });
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------
// v4: "tap window" kernel for stride-1 3x3 convs on small maps (16x16, 32x32 with many channels)
//
// The per-tap kernel (v1) fetches the activation tile once per filter tap and its ring can only keep a fixed number
// of bytes in flight, so it ends up bound by the L2 -> SM traffic (96 channels, fp16x2: 756 KB per 128-pixel tile,
// ~28 B / cycle / SM = two thirds of the measured L2 cap).  Here the output tile is an 8 x 16-pixel block of one
// image; for each kc-channel chunk the CTA loads the block PLUS its halo once (one TMA box of 10 x 18 pixels per
// plane; out-of-image pixels zero-filled = the conv padding) into an A ring slot and runs all nine taps from it:
// tap (r, s) is the same buffer with the descriptor start shifted by (r * 10 + s) rows and the 8-row groups one
// window row (10 pixels) apart (stride byte offset = 10 * SW bytes, see PersistParams::blk).  Only the weight tile of
// (chunk, tap) streams through the B ring.  A traffic drops 9x -> 2.5x the tile (halo), e.g. 393 KB instead of
// 756 KB per tile for 96 channels.  fp16x2 operands as in the per-tap kernel's fat stages: an A slot holds the
// x_hi and the x_lo box, a B stage the row-stacked [w_hi | w_lo] tile.
// ---------------------------------------------------------------------------
constexpr int kTwMaxA = 4, kTwMaxB = 16;
constexpr int kTwWp = 10, kTwHp = 18;        // window = (8 + 2) x (16 + 2) pixels

struct TapWinParams {
  int B, H, W, Cout_p, Cout, Cin_p, relu;
  int tiles_x, tiles_y;      // 8 x 16 blocks per image
  int n_tile, kc, kchunks;   // UMMA N, channels per chunk (SW / 2), chunks of the logical input channels
  int na, nb;                // A ring slots, B ring stages
  int tap_k;                 // K stride between taps in the weight matrix (dense stacked: Cin_p; plain: padded cin_k)
  uint32_t a_bytes, b_bytes; // one activation box (180 * SW), one plane's weight rows (n_tile * SW)
  uint32_t tmem_cols;
  int split, nh;             // fp16x2 storage; K16 slices per plane (plain fp16: all slices)
  const float* bias;
  const __half* res;
  __half* out;
  float* heatmap;            // head1 extras of the shared epilogue (EpiArgs)
  const float* xs;
  const float* ys;
  int coord_maps;
  unsigned long long* ts;    // optional [grid][8] globaltimer stamps (EGN_TC_TS=1)
  // Staged TMA epilogue (see epi_window_staged): once the last MMA has completed the operand rings are dead, so the
  // residual block arrives by TMA at the start of shared memory ([block][8 x 16 pixels][cb channels], hi blocks then
  // lo blocks), every epilogue thread converts its row in place and the block leaves by TMA tensor stores, which
  // also clip ragged tiles.  The direct epilogue (one uncoalesced 16-byte access per lane and instruction) took
  // 12-13 us per 128 x 96 / 192-channel fp16x2 tile -- as long as the MMA phase.
  int staged, cb, nblk_plane;
  uint32_t blk_bytes;
  int keep_a;    // 3-MMA form: the second x_hi MMA takes the A tile from the collector (umma_f16_keep_a)
  // Persistent mode (tiles of <= 256 TMEM columns: two accumulator sets fit): one CTA per SM walks over tiles
  // blockIdx.x, += gridDim.x; the rings run across tile boundaries, the MMAs of tile t+1 overlap the epilogue of tile t
  // (acc_sets = 2) and the staging buffer is a region of its own behind the rings (stage_off != 0, stage_bytes).
  int total_tiles, acc_sets;
  uint32_t stage_off, stage_bytes;
  int grouped;   // fp16x2: issue the MMAs of a stage grouped by accumulator (all [H | L], then all L)
  // N fold (persistent mode): layers wider than 128 channels run as n_fold parts of n_tile channels each, a work unit
  // is (tile, part) -- so that 192-channel layers get two accumulator sets, the cross-tile rings and no wave
  // quantisation (512 tiles on 148 SMs) like the 96-channel ones; the parts of a tile re-load its windows (L2 hits)
  int n_fold;
};

// kPair: the two CTAs of a cluster work as a pair (tcgen05 cta_group::2, M = 256): each CTA loads the windows of its
//   own tile (tiles 2i and 2i + 1) and HALF of the rows of every weight tile -- [w_hi rows n0 + r n/2 .. | w_lo rows ..]
//   at identical shared-memory offsets -- and the leader's MMA thread issues one M = 256 MMA for both.  The weight
//   stream, which every 128-pixel tile has to pull through L2 in full (324 KB for 96 channels, 1.3 MB for 192: at the
//   MMA rate that is ~40 B / cycle / SM, the measured L2 cap), halves per SM.  Barrier protocol as in
//   conv_persist_kernel: a_full / b_full / acc_empty live in the LEADER (the peer's TMA transactions and epilogue
//   arrivals are sent there); a_empty / b_empty / acc_full are signalled in both CTAs by multicast tcgen05.commit.
//   fp16x2 uses the three-MMA form (x_hi w_hi -> H, x_hi w_lo -> L, x_lo w_hi -> L, N = n_tile each): the stacked
//   full-width form would need w_hi in both CTAs at the offset where the peer keeps w_lo.
template <int SW, bool kPair>
__global__ void __launch_bounds__(kTcThreads, 2)
conv_tapwin_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const __grid_constant__ CUtensorMap map_res, const __grid_constant__ CUtensorMap map_out,
                   const TapWinParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_plane = ((uint32_t)(kTwWp * kTwHp) * SW + 1023u) & ~1023u;
  const uint32_t a_slot = p.split ? 2u * a_plane : a_plane;
  const uint32_t b_rows = (uint32_t)(kPair ? p.n_tile / 2 : p.n_tile);       // weight rows per plane held by this CTA
  const uint32_t b_stage = (b_rows * SW * (p.split ? 2u : 1u) + 1023u) & ~1023u;
  const uint32_t crank = kPair ? cluster_ctarank() : 0u;                    // 0 = leader of the pair
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + (size_t)p.na * a_slot;
  uint8_t* smem_stage = smem + p.stage_off;        // stage_off = 0: on top of the (then dead) rings, one tile per CTA
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + (size_t)p.nb * b_stage + p.stage_bytes);
  uint64_t* a_empty = a_full + kTwMaxA;
  uint64_t* b_full = a_empty + kTwMaxA;
  uint64_t* b_empty = b_full + kTwMaxB;
  uint64_t* acc_full = b_empty + kTwMaxB;          // [2]
  uint64_t* acc_empty = acc_full + 2;              // [2]
  uint64_t* res_full = acc_empty + 2;
  uint64_t* stage_free = res_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stage_free + 1);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);          // [n_tile]; offset (2 * 4 + 2 * 16 + 6) * 8 + 16 = 384: 16-byte aligned

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int n_base = blockIdx.y * p.n_tile * p.n_fold;
  // work units of this CTA: blockIdx.x, += gridDim.x (one per CTA unless persistent).  A unit is (tile, N part); pairs:
  // (pair of tiles 2q + rank, N part) -- both CTAs of a pair work on the same part, they share its weight rows (an odd
  // last tile has a past-the-end partner: loads zero-filled, stores clipped)
  const int n_units = (kPair ? (p.total_tiles + 1) / 2 : p.total_tiles) * p.n_fold;
  const int unit0 = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int my_tiles = (n_units - unit0 + unit_step - 1) / unit_step;
  const bool dedicated = p.stage_off != 0;         // staging buffer of its own (persistent mode)
  auto part_of = [&](int it) { return (unit0 + it * unit_step) % p.n_fold; };
  auto n0_of = [&](int it) { return n_base + part_of(it) * p.n_tile; };
  auto tile_xyb = [&](int it, int& x0, int& y0, int& b0) {
    int tile = (unit0 + it * unit_step) / p.n_fold;
    if (kPair) tile = 2 * tile + (int)crank;
    const int tx = tile % p.tiles_x;
    tile /= p.tiles_x;
    x0 = tx * 8;
    y0 = (tile % p.tiles_y) * 16;
    b0 = tile / p.tiles_y;
  };
  const uint32_t set_cols = (uint32_t)((p.split ? 2 : 1) * p.n_tile);      // TMEM columns of one accumulator set
  if (threadIdx.x == 0) EGN_TS(0);
  if (p.staged)
    for (int i = threadIdx.x; i < p.n_tile * p.n_fold; i += (int)blockDim.x) s_bias[i] = __ldg(p.bias + n_base + i);

  if (threadIdx.x == 0) {
    pdl_trigger();
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (p.staged) {
      tma_prefetch_desc(&map_res);
      tma_prefetch_desc(&map_out);
    }
    for (int i = 0; i < kTwMaxA; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kTwMaxB; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kPair ? 8 : 4);     // one arrival per epilogue warp (of both CTAs)
    }
    mbar_init(res_full, 1);
    mbar_init(stage_free, 1);
    fence_barrier_init();
  }
  if constexpr (kPair) {
    cluster_sync_all();                            // both CTAs' barriers initialised before any remote signal
    if (warp == 1) tmem_alloc_pair(tmem_slot, p.tmem_cols);
  } else {
    if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();
  if (threadIdx.x == 0) EGN_TS(1);

  if (warp == 0) {
    // ===================== TMA producer: A chunk windows + the weight tile of every (chunk, tap) =====================
    // q numbers the (tile, chunk) pairs of this CTA: the A ring runs across tile boundaries
    const int Q = my_tiles * p.kchunks;
    auto load_a = [&](int q) {
      const int slot = q % p.na, c = q % p.kchunks;
      int x0, y0, b0;
      tile_xyb(q / p.kchunks, x0, y0, b0);
      mbar_wait(&a_empty[slot], (uint32_t)(((q / p.na) & 1) ^ 1));
      if (elect_one()) {
        uint8_t* sa = smem_a + (size_t)slot * a_slot;
        if constexpr (kPair) {
          // both CTAs' windows are accounted for on the leader's barrier
          if (crank == 0) mbar_expect_tx(&a_full[slot], 2u * (p.split ? 2u : 1u) * p.a_bytes);
          const uint32_t bar = mapa_rank(&a_full[slot], 0);
          tma_load_4d_pair(sa, &map_a, bar, c * p.kc, x0 - 1, y0 - 1, b0);
          if (p.split) tma_load_4d_pair(sa + a_plane, &map_a, bar, p.Cin_p + c * p.kc, x0 - 1, y0 - 1, b0);
        } else {
          mbar_expect_tx(&a_full[slot], (p.split ? 2u : 1u) * p.a_bytes);
          tma_load_4d(sa, &map_a, &a_full[slot], c * p.kc, x0 - 1, y0 - 1, b0);
          if (p.split) tma_load_4d(sa + a_plane, &map_a, &a_full[slot], p.Cin_p + c * p.kc, x0 - 1, y0 - 1, b0);
        }
      }
    };
    // na - 1 chunk windows in flight ahead of the one being multiplied.  The refill is issued at tap 3: the slot was
    // released by the previous chunk's last MMAs, whose commit has certainly arrived by then -- waiting for it at
    // tap 0 / 1 blocks this warp and starves the weight ring (96ch@32x32, persistent: 139 -> 132 us).
    const int D = p.na > 1 ? p.na - 1 : 1;
    for (int q = 0; q < D && q < Q; ++q) load_a(q);
    uint32_t stage = 0, phase = 0;
    for (int q = 0; q < Q; ++q) {
      const int c = q % p.kchunks;
      const int n0 = n0_of(q / p.kchunks);
      for (int tap = 0; tap < 9; ++tap) {
        if (p.na > 1 && tap == 3 && q + D < Q) load_a(q + D);
        mbar_wait(&b_empty[stage], phase ^ 1u);
        if (elect_one()) {
          uint8_t* sb = smem_b + (size_t)stage * b_stage;
          if constexpr (kPair) {
            // this CTA's half of the rows of each plane (p.b_bytes = bytes of that half)
            if (crank == 0) mbar_expect_tx(&b_full[stage], 2u * (p.split ? 2u : 1u) * p.b_bytes);
            const uint32_t bar = mapa_rank(&b_full[stage], 0);
            const int nr = n0 + (int)crank * (p.n_tile / 2);
            tma_load_2d_pair(sb, &map_b, bar, tap * p.tap_k + c * p.kc, nr);
            if (p.split) tma_load_2d_pair(sb + p.b_bytes, &map_b, bar, tap * p.tap_k + c * p.kc, p.Cout_p + nr);
          } else {
            mbar_expect_tx(&b_full[stage], (p.split ? 2u : 1u) * p.b_bytes);
            tma_load_2d(sb, &map_b, &b_full[stage], tap * p.tap_k + c * p.kc, n0);
            if (p.split) tma_load_2d(sb + p.b_bytes, &map_b, &b_full[stage], tap * p.tap_k + c * p.kc, p.Cout_p + n0);
          }
        }
        if (++stage == (uint32_t)p.nb) {
          stage = 0;
          phase ^= 1u;
        }
      }
      // a single window slot is refilled only AFTER the last weight tile of the chunk has been issued: its release
      // (a_empty) is committed behind the tap-8 MMAs, which wait for that tile -- refilling earlier deadlocks
      if (p.na == 1 && q + 1 < Q) load_a(q + 1);
      if (c == p.kchunks - 1 && p.staged && p.res) {
        // residual block of this tile: into the staging buffer once it is free -- its own buffer: the previous tile's
        // stores have been read out; on top of the rings: every MMA of the (only) tile has completed
        const int it = q / p.kchunks;
        if (dedicated) {
          if (it > 0) mbar_wait(stage_free, (uint32_t)((it - 1) & 1));
        } else {
          mbar_wait(&acc_full[0], 0);
        }
        if (elect_one()) {
          int x0, y0, b0;
          tile_xyb(it, x0, y0, b0);
          const int nblk = (p.split ? 2 : 1) * p.nblk_plane;
          mbar_expect_tx(res_full, (uint32_t)nblk * (uint32_t)(128 * p.cb * 2));
          for (int k = 0; k < nblk; ++k) {
            const int ch = k < p.nblk_plane ? n0 + k * p.cb : p.Cout_p + n0 + (k - p.nblk_plane) * p.cb;
            tma_load_4d(smem_stage + (size_t)k * p.blk_bytes, &map_res, res_full, ch, x0, y0, b0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (!kPair || crank == 0) {
    // ===================== MMA issuer (pairs: the leader issues for both CTAs) =====================
    const uint32_t idesc = (1u << 4) | ((uint32_t)(p.n_tile >> 3) << 17) | (((kPair ? 256u : 128u) >> 4) << 24);
    const uint32_t idesc_w = (1u << 4) | ((uint32_t)(p.n_tile >> 2) << 17) | ((128u >> 4) << 24);   // N = 2 n_tile
    const bool wide = !kPair && p.split && 2 * p.n_tile <= 256;
    const uint64_t desc_hi = make_smem_desc(0, SW);
    auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t id, uint32_t acc) {
      if constexpr (kPair) umma_f16_pair(d, ad, bd, id, acc);
      else umma_f16(d, ad, bd, id, acc);
    };
    auto commit = [&](uint64_t* bar) {
      if constexpr (kPair) umma_commit_pair(bar);
      else umma_commit(bar);
    };
    // A: 8-row groups one window row apart
    const uint64_t desc_hi_a = (desc_hi & ~((uint64_t)0x3FFF << 32)) | ((uint64_t)(((uint32_t)kTwWp * SW) >> 4) << 32);
    const uint32_t a_addr0 = smem_u32(smem_a), b_addr0 = smem_u32(smem_b);
    uint32_t stage = 0, phase = 0;
    int q = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const int set = p.acc_sets == 2 ? (it & 1) : 0;
      const uint32_t use = (uint32_t)(p.acc_sets == 2 ? it >> 1 : it);
      mbar_wait(&acc_empty[set], (use & 1u) ^ 1u);         // the epilogue has drained this accumulator set
      tc_fence_after();
      const uint32_t d_h = tmem_base + (uint32_t)set * set_cols, d_l = d_h + (uint32_t)p.n_tile;
      for (int c = 0; c < p.kchunks; ++c, ++q) {
        const int slot = q % p.na;
        mbar_wait(&a_full[slot], (uint32_t)((q / p.na) & 1));
        tc_fence_after();
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(&b_full[stage], phase);
          tc_fence_after();
          if (it == 0 && c == 0 && tap == 0 && lane == 0) EGN_TS(2);
          const int r = tap / 3, qq = tap - 3 * r;
          const uint64_t adesc = desc_hi_a | (uint64_t)(((a_addr0 + (uint32_t)slot * a_slot + (uint32_t)(r * kTwWp + qq) * SW) & 0x3FFFFu) >> 4);
          const uint64_t bdesc = desc_hi | (uint64_t)(((b_addr0 + stage * b_stage) & 0x3FFFFu) >> 4);
          if (elect_one()) {
            const uint64_t adesc_lo = adesc + (uint64_t)(a_plane >> 4);
            const uint64_t bdesc_lo = bdesc + (uint64_t)(p.b_bytes >> 4);
            if (p.split && p.grouped) {
              // same-accumulator runs: alternating the destination between consecutive MMAs costs ~21 cycles per
              // switch (profiles/r01k_umma_rate.log: N = 192 96 -> 118 cycles with two accumulators)
#pragma unroll
              for (int k = 0; k < SW / 32; ++k) {
                if (c * (SW / 32) + k >= p.nh) continue;
                mma(d_h, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), wide ? idesc_w : idesc, (c | tap | k) ? 1u : 0u);
              }
              if (!wide) {
#pragma unroll
                for (int k = 0; k < SW / 32; ++k) {
                  if (c * (SW / 32) + k >= p.nh) continue;
                  mma(d_l, adesc + (uint64_t)(2 * k), bdesc_lo + (uint64_t)(2 * k), idesc, (c | tap | k) ? 1u : 0u);
                }
              }
#pragma unroll
              for (int k = 0; k < SW / 32; ++k) {
                if (c * (SW / 32) + k >= p.nh) continue;
                mma(d_l, adesc_lo + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
              }
            } else
#pragma unroll
            for (int k = 0; k < SW / 32; ++k) {
              if (c * (SW / 32) + k >= p.nh) continue;               // padding slice of the last chunk
              const uint32_t acc = (c | tap | k) ? 1u : 0u;          // the first MMA of a tile initialises the accumulators
              const uint64_t ko = (uint64_t)(2 * k);
              if (!p.split) {
                mma(d_h, adesc + ko, bdesc + ko, idesc, acc);
              } else if (wide) {
                mma(d_h, adesc + ko, bdesc + ko, idesc_w, acc);                   // [H | L] += x_hi [w_hi | w_lo]
                mma(d_l, adesc_lo + ko, bdesc + ko, idesc, 1u);                   // L += x_lo w_hi
              } else {
                if (p.keep_a && !kPair) {
                  umma_f16_keep_a(d_h, adesc + ko, bdesc + ko, idesc, acc);            // H += x_hi w_hi
                  umma_f16_reuse_a(d_l, adesc + ko, bdesc_lo + ko, idesc, acc);        // L += x_hi w_lo
                } else {
                  mma(d_h, adesc + ko, bdesc + ko, idesc, acc);                   // H += x_hi w_hi
                  mma(d_l, adesc + ko, bdesc_lo + ko, idesc, acc);                // L += x_hi w_lo
                }
                mma(d_l, adesc_lo + ko, bdesc + ko, idesc, 1u);                   // L += x_lo w_hi
              }
            }
            commit(&b_empty[stage]);
            if (tap == 8) commit(&a_empty[slot]);
          }
          if (++stage == (uint32_t)p.nb) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      if (elect_one()) commit(&acc_full[set]);
      if (it == 0 && lane == 0) EGN_TS(3);
    }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;     // accumulator row -> pixel (row & 7, row >> 3) of the block
    const int nblk = (p.split ? 2 : 1) * p.nblk_plane;
    for (int it = 0; it < my_tiles; ++it) {
      const int set = p.acc_sets == 2 ? (it & 1) : 0;
      const uint32_t use = (uint32_t)(p.acc_sets == 2 ? it >> 1 : it);
      int x0, y0, b0;
      tile_xyb(it, x0, y0, b0);
      const int n0 = n0_of(it);
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)set * set_cols;
      if (p.staged) {
        const EpiStage es{smem_u32(smem_stage), smem_u32(s_bias + part_of(it) * p.n_tile), p.blk_bytes, p.cb, p.res != nullptr, p.relu, 0,
                          p.split ? (uint32_t)p.nblk_plane * p.blk_bytes : 0u};
        if (p.res) mbar_wait(res_full, (uint32_t)(it & 1));
        else if (dedicated && it > 0) mbar_wait(stage_free, (uint32_t)((it - 1) & 1));      // previous stores read out
        mbar_wait(&acc_full[set], use & 1u);
        tc_fence_after();
        if (it == 0 && (threadIdx.x & 127) == 64) EGN_TS(4);
        epi_window_staged<1>(es, tbase, 1, p.n_tile, 0, [&](int) {
          StageRow sr;
          sr.valid = true;                     // pixels past the image are clipped by the TMA store
          sr.srow = (uint32_t)row;             // staging row = (row >> 3) * 8 + (row & 7): the block is row-major
          return sr;
        }, p.split ? 2 : 1, p.split);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {                                      // accumulator set may be overwritten
          if constexpr (kPair) mbar_arrive_cluster(mapa_rank(&acc_empty[set], 0));
          else mbar_arrive(&acc_empty[set]);
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");        // the four epilogue warps
        if (warp == 2 && elect_one()) {
          for (int k = 0; k < nblk; ++k) {
            const int ch = k < p.nblk_plane ? n0 + k * p.cb : p.Cout_p + n0 + (k - p.nblk_plane) * p.cb;
            tma_store_4d(&map_out, smem_stage + (size_t)k * p.blk_bytes, ch, x0, y0, b0);
          }
          bulk_commit();
          bulk_wait_read0();                                  // the buffer has been read: it may be refilled
          if (dedicated) mbar_arrive(stage_free);
        }
      } else {
        EpiRow er;
        er.b = b0;
        er.oh = y0 + (row >> 3);
        er.ow = x0 + (row & 7);
        er.valid = er.b < p.B && er.oh < p.H && er.ow < p.W;
        er.pix = ((size_t)er.b * p.H + er.oh) * p.W + er.ow;
        EpiArgs e{p.bias, p.res, p.out, p.heatmap, p.xs, p.ys, p.coord_maps, p.relu, p.Cout, p.Cout_p, p.H, p.W, 0,
                  (p.ts && it == 0) ? p.ts + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr, p.split, (uint32_t)p.n_tile,
                  0u};
        epi_run(e, tbase, 1, p.n_tile, n0, &acc_full[set], [&](int) { return er; }, use & 1u);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kPair) mbar_arrive_cluster(mapa_rank(&acc_empty[set], 0));
          else mbar_arrive(&acc_empty[set]);
        }
      }
      if (it == 0 && (threadIdx.x & 127) == 64) EGN_TS(5);
    }
    if (p.staged && warp == 2 && elect_one()) bulk_wait0();   // all output writes complete before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();                    // the peer's smem / barriers stay valid until both are done
  if (warp == 1) {
    tc_fence_after();
    if constexpr (kPair) tmem_dealloc_pair(tmem_base, p.tmem_cols);
    else tmem_dealloc(tmem_base, p.tmem_cols);
    if (lane == 0) EGN_TS(6);
  }
}

// ---------------------------------------------------------------------------
// v2: "window run" kernel for stride-1 convs (3x3 pad 1 and 1x1)
//
// v1 re-fetches the activation tile from L2 once per filter tap (9x the bytes).  Here a CTA loads
// ONE zero-padded window of the input -- TBW images x (THW+2) rows x (W+2) pixels x Cin, a single
// TMA box per 64-channel chunk whose out-of-image parts are zero-filled by the TMA unit -- and keeps
// it in shared memory.  In the flattened window (row pitch Wp = W+2 pixels) the operand of tap
// (r,s) for output run position m is simply window row m + r*Wp + s, so all nine taps are the SAME
// shared-memory buffer addressed through UMMA descriptors whose start address is shifted by
// (r*Wp + s) rows.  (Swizzle is applied on absolute smem address bits, so a row shift that is not
// a multiple of the 8-row atom is legal with base_offset = 0: verified on hardware by
// egn_debug_umma_probe, profiles/r01_umma_row_offset_probe.json.)  The two pixels per image row
// that fall on the window border produce junk accumulator rows which the epilogue skips.
//
// Loop order is (tap, chunk) outer / M-tile inner with all T accumulators resident in TMEM, so
// every weight tile is fetched once per CTA and streamed through a small ring.
// ---------------------------------------------------------------------------
constexpr int kMaxChunks = 6;   // Cin_p <= 384
constexpr int kMaxChunksPersist = 12;   // v3 has no per-chunk barriers; fp16x2 storage doubles the physical channels
constexpr int kMaxBStages = 6;

struct RunParams {
  int B, H, W, Cout_p, Cout, Cin_p;
  int taps, relu;
  int halo, Wp, Hw;          // halo = 1 for 3x3, 0 for 1x1; window = TBW x Hw x Wp pixels
  int THW, TBW, win_per_img;
  int lead;                  // halo * (Wp + 1): window row of run position 0
  int m_run, T;              // run length (rows) and number of 128-row M tiles
  int rows_alloc;            // smem rows per chunk buffer
  int n_tile, kchunks, cin_k, b_stages;
  uint32_t a_bytes, b_bytes, tmem_cols;
  const float* bias;
  const __half* res;
  __half* out;
  float* heatmap;
  const float* xs;
  const float* ys;
  int coord_maps;
  float inv_wp, inv_img;     // 1/Wp, 1/(Hw*Wp)
  int img_rows;              // Hw * Wp
  int dbg;   // tuning experiments (EGN_TC_DBG): 1 no stores, 2 no residual, 4 no MMA, 8 no A load, 16 no B loads
  unsigned long long* ts;  // optional [grid][8] globaltimer stamps (EGN_TC_TS=1)
  // fp16x2 split storage in the window-run kernel (v2): Cin_p counts the physical channels ([x_hi | x_lo] planes of
  // run_nh K16 slices each), the streamed weight matrix holds [w_hi | w_lo] per tap, and weight slice sb pairs with
  // activation slices as in PersistParams (hi*hi -> accumulator H, cross terms -> accumulator L = H + T * n_tile)
  int run_split, run_nh;
};

__global__ void __launch_bounds__(kTcThreads, 2)
conv_run_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const RunParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_chunk = (uint32_t)p.rows_alloc * 128u;              // multiple of 1024
  const uint32_t b_stage = ((uint32_t)p.n_tile * 128u + 1023u) & ~1023u;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + (size_t)p.kchunks * a_chunk;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + (size_t)p.b_stages * b_stage);
  uint64_t* b_full = a_full + kMaxChunks;
  uint64_t* b_empty = b_full + kMaxBStages;
  uint64_t* tmem_full_bar = b_empty + kMaxBStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 6);        // [n_tile], 16-byte aligned (offset 176)

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) EGN_TS(0);
  const int win = blockIdx.x % p.win_per_img;
  const int bg = blockIdx.x / p.win_per_img;
  const int h0 = win * p.THW, b0 = bg * p.TBW;
  const int n0 = blockIdx.y * p.n_tile;
  if (threadIdx.x >= 64)
    for (int i = threadIdx.x - 64; i < p.n_tile; i += kTcThreads - 64) s_bias[i] = __ldg(p.bias + n0 + i);

  if (threadIdx.x == 0) {
    pdl_trigger();
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int c = 0; c < p.kchunks; ++c) mbar_init(&a_full[c], 1);
    for (int s = 0; s < p.b_stages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();   // everything above is independent of the previous kernel (common.h, PDL rules)
  if (threadIdx.x == 0) EGN_TS(1);
  const int ntap = p.halo ? 3 : 1;                 // taps per axis
  const int last_ksteps = (p.Cin_p - (p.kchunks - 1) * 64 + 15) >> 4;   // K16 slices holding real channels

  // Producer and MMA warps run their loops warp-uniformly (all 32 lanes wait on the barriers and
  // compute identical addresses / descriptors, which therefore live in uniform registers); only the
  // TMA / tcgen05 instructions themselves are predicated to lane 0.  Keeping the address math out of
  // a divergent region matters: otherwise every UTCHMMA is wrapped in a waterfall loop.
  if (warp == 0) {
    // ===================== TMA producer =====================
    for (int c = 0; c < p.kchunks && !(p.dbg & 8); ++c) {
      if (elect_one()) {
        mbar_expect_tx(&a_full[c], p.a_bytes);
        tma_load_4d(smem_a + (size_t)c * a_chunk, &map_a, &a_full[c], c * 64, -p.halo, h0 - p.halo, b0);
      }
    }
    if (!(p.dbg & 16)) {
      uint32_t stage = 0, phase = 0;
      int kcoord = 0;                              // tap * cin_k + c * 64, advanced incrementally
      for (int tap = 0; tap < p.taps; ++tap) {
        for (int c = 0; c < p.kchunks; ++c, kcoord += 64) {
          mbar_wait(&b_empty[stage], phase ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(&b_full[stage], p.b_bytes);
            tma_load_2d(smem_b + stage * b_stage, &map_b, &b_full[stage], kcoord, n0);
          }
          if (++stage == (uint32_t)p.b_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = (1u << 4) | ((uint32_t)(p.n_tile >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t desc_hi = make_smem_desc(0, 128);          // everything but the start-address field
    const uint32_t a_addr0 = smem_u32(smem_a), b_addr0 = smem_u32(smem_b);
    uint32_t stage = 0, phase = 0, accumulate = 0;
    bool first_tap = true;
    if (p.run_split && !(p.dbg & 8)) {
      // cross pairing reads activation chunks out of order: the whole window must have landed
      for (int c = 0; c < p.kchunks; ++c) mbar_wait(&a_full[c], 0);
      first_tap = false;
    }
    for (int r = 0; r < ntap; ++r) {
      for (int q = 0; q < ntap; ++q) {
        const uint32_t shift = (uint32_t)(r * p.Wp + q) * 128u;   // row shift of this tap inside the window
        const bool tap0 = r == 0 && q == 0;
        for (int c = 0; c < p.kchunks; ++c) {
          if (first_tap && !(p.dbg & 8)) {
            mbar_wait(&a_full[c], 0);
            if (c == 0 && lane == 0) EGN_TS(2);
          }
          if (!(p.dbg & 16)) mbar_wait(&b_full[stage], phase);
          tc_fence_after();
          const uint64_t ad0 = desc_hi | (uint64_t)(((a_addr0 + (uint32_t)c * a_chunk + shift) & 0x3FFFFu) >> 4);
          const uint64_t bd0 = desc_hi | (uint64_t)(((b_addr0 + stage * b_stage) & 0x3FFFFu) >> 4);
          const int ksteps = (c == p.kchunks - 1) ? last_ksteps : 4;
          if (!(p.dbg & 4) && elect_one()) {
            if (!p.run_split) {
              for (int k = 0; k < ksteps; ++k) {
                uint64_t ad = ad0 + (uint64_t)(2 * k);
                uint32_t d = tmem_base;
                for (int t = 0; t < p.T; ++t, ad += (128u * 128u) >> 4, d += (uint32_t)p.n_tile)
                  umma_f16(d, ad, bd0 + (uint64_t)(2 * k), idesc, accumulate | (uint32_t)(k > 0));
              }
            } else {
              for (int k = 0; k < ksteps; ++k) {
                const int sb = 4 * c + k;                       // weight slice of this tap
                const int npart = sb < p.run_nh ? 2 : 1;
                for (int part = 0; part < npart; ++part) {
                  const int lo = sb < p.run_nh ? part : 1;      // accumulator: 0 = H, 1 = L
                  const int sa = sb < p.run_nh ? sb + part * p.run_nh : sb - p.run_nh;
                  uint64_t ad = desc_hi | (uint64_t)(((a_addr0 + (uint32_t)(sa >> 2) * a_chunk + shift) & 0x3FFFFu) >> 4);
                  ad += (uint64_t)(2 * (sa & 3));
                  uint32_t d = tmem_base + (uint32_t)(lo * p.T * p.n_tile);
                  // H is first written by (tap 0, slice 0, x_hi), L by (tap 0, slice 0, x_lo)
                  const uint32_t acc = (tap0 && sb == 0) ? 0u : 1u;
                  for (int t = 0; t < p.T; ++t, ad += (128u * 128u) >> 4, d += (uint32_t)p.n_tile)
                    umma_f16(d, ad, bd0 + (uint64_t)(2 * k), idesc, acc);
                }
              }
            }
          }
          accumulate = 1u;
          if (elect_one()) umma_commit(&b_empty[stage]);
          if (++stage == (uint32_t)p.b_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        first_tap = false;
      }
    }
    if (elect_one()) {
      umma_commit(tmem_full_bar);
      EGN_TS(3);
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    EpiArgs e{s_bias - n0, p.res, p.out, p.heatmap, p.xs, p.ys, p.coord_maps, p.relu, p.Cout, p.Cout_p, p.H, p.W, p.dbg,
              p.ts ? p.ts + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr, p.run_split,
              (uint32_t)(p.T * p.n_tile), 0u};
    epi_run(e, tmem_base + ((uint32_t)(quarter * 32) << 16), p.T, p.n_tile, n0, tmem_full_bar, [&](int t) {
      // run position -> window position -> (image, row, column); float reciprocals are exact here
      // (positions < 2^16, margins >= 0.5 / pitch)
      const int pos = t * 128 + row + p.lead;
      const int bi = p.TBW == 1 ? 0 : (int)(((float)pos + 0.5f) * p.inv_img);
      const int rem = pos - bi * p.img_rows;
      const int hp = (int)(((float)rem + 0.5f) * p.inv_wp);
      const int wp = rem - hp * p.Wp;
      EpiRow rr;
      rr.b = b0 + bi;
      rr.oh = h0 + hp - p.halo;
      rr.ow = wp - p.halo;
      rr.valid = pos - p.lead < p.m_run && bi < p.TBW && hp >= p.halo && hp < p.Hw - p.halo && wp >= p.halo &&
                 wp < p.Wp - p.halo && rr.b < p.B && rr.oh < p.H;
      rr.pix = ((size_t)rr.b * p.H + rr.oh) * p.W + rr.ow;
      return rr;
    });
    if ((threadIdx.x & 127) == 64) EGN_TS(5);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
    if (lane == 0) EGN_TS(6);
  }
}

// ---------------------------------------------------------------------------
// v3: persistent window-run kernel (one CTA per SM, 12 warps)
//
// Same operand addressing as v2, but the CTA stays resident and walks over windows
// (w = blockIdx.x, += gridDim.x) with every phase of consecutive windows overlapped:
//   warp 0      A producer: double-buffered input windows (a_full / a_empty)
//   warp 1      MMA issuer: accumulates window j into TMEM set j&1 (acc_full / acc_empty)
//   warp 2      B producer: the whole weight matrix once if it fits in smem (w_full), else a ring
//   warps 4-11  epilogue: all eight warps drain the accumulator set of the window just finished
//               (two warps per TMEM lane quarter, items split in a checkerboard)
// so while the tensor pipe works on window j, window j+1 is landing in the other smem slot and the
// accumulators of window j-1 are being converted / stored.
// ---------------------------------------------------------------------------
constexpr int kPersistThreadsHead = 384;   // 4 role warps + 2 x 4 epilogue warps (generic epilogue, 161 registers)
constexpr int kPersistThreads = 640;       // 4 role warps + 4 x 4 epilogue warps (compact epilogue)

struct PersistParams {
  RunParams r;          // geometry / operands as in v2 (r.b_stages = ring depth when streaming)
  int n_windows;        // win_per_img * ceil(B / TBW)
  int b_resident;       // 1: all taps*kchunks weight tiles live in smem for the CTA's lifetime
  int b_rows;           // weight rows per CTA tile: n_tile, or n_tile / 2 in CTA-pair mode (each CTA holds half of N)
  int ksplit;           // accumulators per M tile (K loop dealt round-robin; summed by the epilogue), >= 1
  int pack_tail;        // 1: 3x3 conv whose last 64-channel chunk holds 32 real channels -- the weight matrix keeps
                        //    those tail chunks of taps (2i, 2i+1) side by side in ONE 64-wide tile (resident mode)
  int w_tiles;          // weight tiles resident per CTA: taps * kchunks, or taps * (kchunks-1) + ceil(taps/2) packed
  // staged epilogue (0 = direct global accesses)
  int n_stage;          // staging buffers: 0, 1 or 2
  int cb, nblk;         // channels per staging block (64 / 48), blocks per n_tile
  int rows_stage;       // TBW * THW * W pixels per window
  uint32_t blk_bytes;   // bytes of one block, 1024-aligned
  // fp16x2 split storage (kernels.h): r.Cin_p = 2 * Cin physical channels ([x_hi | x_lo] planes, nh K16 slices each).
  // STACKED weights: a weight tile holds 2 * n_tile rows, [w_hi rows | w_lo rows], over 64 K columns = four K16
  // slices of ONE plane, the (tap, slice) pairs packed densely (slice tap * nh + s: no per-tap padding), and tile t's
  // accumulators sit side by side in TMEM, [H_t | L_t].  Per (tap, slice s) the issuer runs
  //   A = x_hi[s], B = all 2 * n_tile rows, N = 2 * n_tile  ->  H += x_hi w_hi,  L += x_hi w_lo   (one MMA)
  //   A = x_lo[s], B = the first n_tile rows, N = n_tile    ->  L += x_lo w_hi
  // two MMAs per slice instead of three and the x_hi operand is read from shared memory once (an N <= 128 MMA is bound
  // by its operand reads: 32 + N / 4 cycles, so 56 + 44 = 100 cycles per slice at n_tile = 48 instead of 3 x 44).
  // H and L are summed by the epilogue (ksplit = 2, accumulator stride n_tile).  The staging buffer holds nblk_plane
  // hi blocks followed by nblk_plane lo blocks (nblk = 2 * nblk_plane); out / res tensors have 2 * Cout_p channels.
  int split, nh, nblk_plane;
  int a_sw;             // bytes per window row and chunk: 128 (64 channels, SWIZZLE_128B) or 64 (32 channels, SWIZZLE_64B:
                        // 96 physical channels = three exact chunks instead of two with a half-empty second one)
  int a_slots;          // input-window slots in smem: 2 (window j+1 loads while j is multiplied) or 1 (load and MMA
                        // phases of consecutive windows serialise; the epilogue still overlaps -- TMEM stays double buffered)
  // chunk phasing (ONE window slot of 2..4 chunks, fp16x2): the issue table is ordered phase c = every MMA whose
  // activation slice lies in chunk c, and each chunk has its own full / empty barrier (a_full[c] / a_empty[c]).
  // Chunk c of the next window reloads while the other phases run: the single slot behaves like a ring of chunks.
  // phase_end[c] = issue-table index one past phase c.
  int chunk_phase, phase_end[4];
  int grouped;          // fp16x2: inside a phase, issue all full-width MMAs first, then all half-width ones
  // Block-shaped windows (3x3 convs): a window is BW x BH output pixels (BW = 8 * TX, BH = 16 * TY) plus the halo,
  // loaded as ONE TMA box of r.Wp = BW + 2 columns x r.Hw = BH + 2 rows.  An M tile is 8 columns x 16 rows: its 16
  // 8-row operand groups are one WINDOW ROW (r.Wp pixels) apart, which a K-major UMMA descriptor expresses directly
  // as its stride byte offset (r.Wp * 128 B; any multiple of 16 B is legal, verified by tools/probes/
  // umma_group_stride.py).  Every accumulator row is a real output pixel: no junk rows at the image borders (the
  // flattened full-width run wastes 2 of every W + 2 rows plus the tail of the last tile: 73 % useful rows for
  // 64 x 64 maps at T = 2), and the halo is loaded for 324 instead of 330 pixels per 256 (192) outputs.
  int blk, BW, BH, TX, wins_x;
};

// kHead: generic epilogue with the head1 extras (fp32 heat-map copy, coordinate maps).
// kPair: the CTAs of a 2-CTA cluster work as a pair (tcgen05 cta_group::2).  Each CTA loads its own window (its
//   128 x T rows of the M = 256 x T tile) and keeps HALF of the output channels' weights resident; the leader's
//   MMA thread issues one N = n_tile MMA for both, so a layer whose weights only fit when split over N runs one
//   pass of N = 96 MMAs instead of two passes of N = 48 (same operand bytes per MMA, twice the work).
//   Barrier protocol: a_full / w_full / acc_empty live in the LEADER (the peer's TMA transactions and epilogue
//   arrivals are sent there); a_empty / acc_full are signalled in both CTAs by multicast tcgen05.commit.
template <bool kHead, bool kPair = false>
__global__ void __launch_bounds__(kHead ? kPersistThreadsHead : kPersistThreads, 1)
conv_persist_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_res, const __grid_constant__ CUtensorMap map_out,
                    const PersistParams pp) {
  const RunParams& p = pp.r;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_chunk = (uint32_t)p.rows_alloc * (uint32_t)pp.a_sw;
  const uint32_t a_slot = a_chunk * (uint32_t)p.kchunks;
  const uint32_t b_stage = ((uint32_t)pp.b_rows * (pp.split ? 256u : 128u) + 1023u) & ~1023u;   // stacked: [w_hi | w_lo] rows
  const int w_tiles = pp.w_tiles;
  const int b_slots = pp.b_resident ? w_tiles : p.b_stages;
  const uint32_t crank = kPair ? cluster_ctarank() : 0u;          // 0 = leader of the pair
  // window of iteration j: w0 + j * gridDim.x + w_off (pair: consecutive windows 2q, 2q+1 per iteration; the odd
  // one may lie past the end -- its loads are zero-filled by TMA and its rows are invalid in the epilogue)
  const int w_off = kPair ? (int)crank : 0;
  const int w0 = (int)blockIdx.x - w_off;
  uint8_t* smem_a = smem;                                   // 2 slots
  uint8_t* smem_b = smem + (size_t)pp.a_slots * a_slot;
  uint8_t* smem_stage = smem_b + (size_t)b_slots * b_stage;                  // n_stage x nblk x blk_bytes (1024-aligned)
  const uint32_t stage_bytes = (uint32_t)pp.nblk * pp.blk_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_stage + (size_t)pp.n_stage * stage_bytes);
  uint64_t* a_empty = a_full + 4;                   // [4]: window slots, or the chunks of a chunk-phased slot
  uint64_t* acc_full = a_empty + 4;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* w_full = acc_empty + 2;
  uint64_t* b_full = w_full + 1;
  uint64_t* b_empty = b_full + kMaxBStages;
  uint64_t* stage_full = b_empty + kMaxBStages;     // [2] residual landed / buffer free for the epilogue
  uint64_t* staged = stage_full + 2;                // [2] epilogue finished writing the buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(staged + 2);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 2);    // offset (13 + 12 + 4) * 8 + 8 = 240: 16-byte aligned
  uint4* s_issue = reinterpret_cast<uint4*>(s_bias + p.n_tile);   // [n_mma] issue table (n_tile % 16 == 0)

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * p.n_tile;
  constexpr int kParts = kHead ? 2 : 4;              // epilogue warps per TMEM lane quarter
  for (int i = threadIdx.x; i < p.n_tile; i += (int)blockDim.x) s_bias[i] = __ldg(p.bias + n0 + i);
  // Issue table: one entry per tcgen05.mma of a window, in issue order (tap, 64-channel chunk, K16
  // slice, M tile).  The single issuing thread then runs a flat loop of {LDS.128, 3 adds, MMA}; the
  // nested-loop form spent ~80 SM cycles of scalar work per MMA, twice what the tensor pipe needs
  // for a 128x48x16 tile (profiles/r01_umma_rate.json).
  //   x: A descriptor start-address offset from the window slot (>>4)   y: same for the weights
  //   z: TMEM column offset of the M tile   w: bit0 accumulate, bit1 first / bit2 last MMA of a weight tile
  {
    const int ktot = (p.Cin_p + 15) >> 4;                      // K16 slices per tap (both planes in split mode)
    const int groups = pp.split ? 2 * pp.nh : ktot;            // MMA groups (of T tiles) per tap
    const int n_mma = p.taps * groups * p.T;
    const int ntap_w = p.halo ? 3 : 1;
    const int spc = pp.a_sw >> 5;                              // K16 slices per window chunk
    const uint32_t b_stage_t = b_stage;
    // activation slice of group g of a tap (split storage, stacked weights: g even = x_hi[g/2], odd = x_lo[g/2])
    auto slice_of = [&](int g) { return pp.split ? (g >> 1) + (g & 1) * pp.nh : g; };
    for (int i = threadIdx.x; i < n_mma; i += (int)blockDim.x) {
      int t = i % p.T, kk = i / p.T;
      int tap = kk / groups, g = kk - tap * groups;
      if (pp.chunk_phase || (pp.split && pp.grouped)) {
        // issue position -> (phase = chunk, [accumulator run], tap, rank inside the phase) -> group g.  Without chunk
        // phasing the whole window is one phase.  Grouped (fp16x2): inside a phase all full-width MMAs ([H | L]) come
        // first, then all half-width ones (L) -- alternating the destination costs ~21 cycles per switch.
        int ph = 0;
        if (pp.chunk_phase)
          while (i >= pp.phase_end[ph]) ++ph;
        const int base = (pp.chunk_phase && ph) ? pp.phase_end[ph - 1] : 0;
        auto in_phase = [&](int gg) { return !pp.chunk_phase || slice_of(gg) / spc == ph; };
        int per[2] = {0, 0};                              // groups of a tap in this phase: full-width, half-width
        for (int gg = 0; gg < groups; ++gg)
          if (in_phase(gg)) ++per[pp.split ? (gg & 1) : 0];
        int j = i - base, want = -1;                      // want: 0 / 1 = accumulator run, -1 = tap-major order
        int per_run = per[0] + per[1];
        if (pp.split && pp.grouped) {
          want = j < p.taps * per[0] * p.T ? 0 : 1;
          if (want) j -= p.taps * per[0] * p.T;
          per_run = per[want];
        }
        t = j % p.T;
        const int kq = j / p.T;
        tap = kq / per_run;
        int rank = kq - tap * per_run;
        for (g = 0; g < groups; ++g)
          if (in_phase(g) && (want < 0 || (g & 1) == want) && rank-- == 0) break;
      }
      // A slice sa, weight slice sb; lo = 1: the half-width MMA x_lo * w_hi into L
      int sa = g, sb = g, lo = 0;
      if (pp.split) {
        sb = g >> 1;
        lo = g & 1;
        sa = sb + lo * pp.nh;
      }
      const int ca = sa / spc, ka = sa - ca * spc;
      const int c = sb >> 2, k = sb & 3;
      const int r = tap / ntap_w, q = tap - r * ntap_w;
      const int ksteps_c = (c == p.kchunks - 1) ? ktot - 4 * c : 4;
      uint4 e;
      // first operand row of (tap, tile): flattened run -> tap shift + 128 rows per tile; block windows -> tile
      // (tx, ty) starts 8 * tx columns / 16 * ty rows into the window
      const uint32_t row0 = pp.blk ? (uint32_t)((r + 16 * (t / pp.TX)) * p.Wp + q + 8 * (t % pp.TX))
                                   : (uint32_t)(r * p.Wp + q + t * 128);
      e.x = (((uint32_t)ca * a_chunk + row0 * (uint32_t)pp.a_sw) >> 4) + 2u * (uint32_t)ka;
      // weight tile and K16 slot inside it: (tap, chunk) tiles in order, or -- packed -- the full chunks first and
      // then one tile per pair of taps holding both 32-channel tails (slots 0-1: even tap, 2-3: odd tap);
      // stacked (split): the slices of all taps packed densely, four per tile
      uint32_t wt = (uint32_t)(tap * p.kchunks + c), ks = (uint32_t)k;
      if (pp.split) {
        const int ws = tap * pp.nh + sb;
        wt = (uint32_t)(ws >> 2);
        ks = (uint32_t)(ws & 3);
      } else if (pp.pack_tail) {
        if (c < p.kchunks - 1) {
          wt = (uint32_t)(tap * (p.kchunks - 1) + c);
        } else {
          wt = (uint32_t)(p.taps * (p.kchunks - 1) + (tap >> 1));
          ks = (uint32_t)((tap & 1) * 2 + k);
        }
      }
      e.y = (pp.b_resident ? (wt * b_stage_t) >> 4 : 0u) + 2u * ks;
      if (pp.split) {
        // tile t: [H_t | L_t]; the full-width MMA starts at H_t, the half-width one at L_t.  The first MMA issued
        // for a tile (a full-width one: every phase order starts with an x_hi slice) initialises both halves.
        e.z = (uint32_t)(t * 2 * p.n_tile + lo * p.n_tile);
        e.w = (i < p.T ? 0u : 1u) | (lo ? 16u : 0u);
      } else {
        e.z = (uint32_t)(((kk % pp.ksplit) * p.T + t) * p.n_tile);
        e.w = kk >= pp.ksplit ? 1u : 0u;
        e.w |= ((k == 0 && t == 0) ? 2u : 0u) | ((k == ksteps_c - 1 && t == p.T - 1) ? 4u : 0u);   // weight-ring tile boundaries
      }
      s_issue[i] = e;
    }
  }
  if (threadIdx.x == 0) {
    pdl_trigger();
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (pp.n_stage) {
      tma_prefetch_desc(&map_res);
      tma_prefetch_desc(&map_out);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], (kPair ? 2 : 1) * 4 * kParts);        // one arrival per epilogue warp (of both CTAs)
      mbar_init(&stage_full[i], 1);
      mbar_init(&staged[i], 4 * kParts);
    }
    mbar_init(w_full, 1);
    for (int i = 0; i < kMaxBStages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    fence_barrier_init();
  }
  if constexpr (kPair) {
    cluster_sync_all();                            // both CTAs' barriers initialised before any remote signal
    if (warp == 1) tmem_alloc_pair(tmem_slot, p.tmem_cols);
  } else {
    if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const int acc_cols = pp.ksplit * p.T * p.n_tile;           // TMEM columns of one accumulator set

  // PDL: the weight producer (warp 2) and the MMA issuer (warp 1) never touch activation memory and start
  // at once -- the resident weights stream in while the previous kernel drains; the roles that read or
  // write activations (A producer, staging DMA, epilogue) first wait for the previous grid to complete.
  if (warp == 0) {
    // ===================== A producer =====================
    pdl_wait();
    int j = 0;
    for (int wb = w0; wb < pp.n_windows; wb += gridDim.x, ++j) {
      const int w = wb + w_off;
      const int slot = pp.a_slots == 2 ? (j & 1) : 0;
      const int bg = w / p.win_per_img, win = w - bg * p.win_per_img;
      // TMA coordinates of the window's first (halo) pixel
      const int wy = pp.blk ? win / pp.wins_x : win, wx = pp.blk ? win - wy * pp.wins_x : 0;
      const int cx = wx * pp.BW - p.halo, cy = (pp.blk ? wy * pp.BH : win * p.THW) - p.halo, cb0 = bg * p.TBW;
      const int cpc = pp.a_sw >> 1;                   // channels per chunk
      // direct epilogue: pull the residual pixels of this window towards L2 now, ~2 windows before they are read
      auto res_prefetch = [&]() {
        if (!p.res || pp.n_stage || (p.dbg & 128) || w >= pp.n_windows) return;
        const int cpitch = pp.split ? 2 * p.Cout_p : p.Cout_p;
        if (pp.blk) {
          // one piece per block row (lane = row)
          const int y = wy * pp.BH + lane, x0 = wx * pp.BW;
          if (lane < pp.BH && y < p.H && x0 < p.W)
            l2_prefetch_bulk(p.res + (((size_t)cb0 * p.H + y) * p.W + x0) * cpitch, (uint32_t)(min(pp.BW, p.W - x0) * cpitch * 2));
        } else if (elect_one()) {
          // the rows of a full-width window are one contiguous NHWC range
          const int h0 = win * p.THW;
          const int rows = p.TBW > 1 ? min(p.TBW, p.B - cb0) * p.H : min(p.THW, p.H - h0);
          l2_prefetch_bulk(p.res + ((size_t)cb0 * p.H + h0) * p.W * cpitch, (uint32_t)(rows * p.W * cpitch * 2));
        }
      };
      if (pp.chunk_phase) {
        // one slot, kchunks chunks, each released by its own phase of the MMA loop
        for (int c = 0; c < p.kchunks; ++c) {
          mbar_wait(&a_empty[c], (uint32_t)((j & 1) ^ 1));
          if (p.ts && j < 8 && lane == 0 && c < 2) p.ts[((size_t)blockIdx.x * 8 + j) * 16 + 10 + c] = gtime();
          if (!(p.dbg & 8) && elect_one()) {
            mbar_expect_tx(&a_full[c], p.a_bytes);
            tma_load_4d(smem_a + (size_t)c * a_chunk, &map_a, &a_full[c], c * cpc, cx, cy, cb0);
          }
        }
        res_prefetch();
        continue;
      }
      mbar_wait(&a_empty[slot], (uint32_t)(((pp.a_slots == 2 ? j >> 1 : j) & 1) ^ 1));
      if (!(p.dbg & 8) && elect_one()) {
        if constexpr (kPair) {
          // both windows of the iteration are accounted for on the leader's barrier
          if (crank == 0) mbar_expect_tx(&a_full[slot], 2u * p.a_bytes * (uint32_t)p.kchunks);
          const uint32_t bar = mapa_rank(&a_full[slot], 0);
          for (int c = 0; c < p.kchunks; ++c)
            tma_load_4d_pair(smem_a + (size_t)slot * a_slot + (size_t)c * a_chunk, &map_a, bar, c * cpc, cx, cy, cb0);
        } else {
          mbar_expect_tx(&a_full[slot], p.a_bytes * (uint32_t)p.kchunks);
          for (int c = 0; c < p.kchunks; ++c)
            tma_load_4d(smem_a + (size_t)slot * a_slot + (size_t)c * a_chunk, &map_a, &a_full[slot], c * cpc, cx, cy, cb0);
        }
      }
      if (!(p.dbg & 8)) res_prefetch();
    }
  } else if (warp == 2) {
    // ===================== B producer =====================
    if (pp.b_resident) {
      if (elect_one()) {
        int kcoord = 0;
        if constexpr (kPair) {
          // this CTA's half of the output channels; both halves are accounted for on the leader's barrier
          if (crank == 0) mbar_expect_tx(w_full, 2u * p.b_bytes * (uint32_t)w_tiles);
          const uint32_t bar = mapa_rank(w_full, 0);
          for (int i = 0; i < w_tiles; ++i, kcoord += 64)
            tma_load_2d_pair(smem_b + (size_t)i * b_stage, &map_b, bar, kcoord, n0 + (int)crank * pp.b_rows);
        } else if (pp.split) {
          // stacked tiles: rows [0, n_tile) = w_hi, [n_tile, 2 n_tile) = w_lo (global rows n0.. and Cout_p + n0..)
          mbar_expect_tx(w_full, 2u * p.b_bytes * (uint32_t)w_tiles);
          for (int i = 0; i < w_tiles; ++i, kcoord += 64) {
            tma_load_2d(smem_b + (size_t)i * b_stage, &map_b, w_full, kcoord, n0);
            tma_load_2d(smem_b + (size_t)i * b_stage + p.b_bytes, &map_b, w_full, kcoord, p.Cout_p + n0);
          }
        } else {
          mbar_expect_tx(w_full, p.b_bytes * (uint32_t)w_tiles);
          for (int i = 0; i < w_tiles; ++i, kcoord += 64)
            tma_load_2d(smem_b + (size_t)i * b_stage, &map_b, w_full, kcoord, n0);
        }
      }
    } else {
      uint32_t stage = 0, phase = 0;
      for (int w = blockIdx.x; w < pp.n_windows; w += gridDim.x) {
        int kcoord = 0;
        for (int i = 0; i < w_tiles; ++i, kcoord += 64) {
          mbar_wait(&b_empty[stage], phase ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(&b_full[stage], p.b_bytes);
            tma_load_2d(smem_b + stage * b_stage, &map_b, &b_full[stage], kcoord, n0);
          }
          if (++stage == (uint32_t)p.b_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // One elected thread runs the whole role (waits, MMAs, commits): no per-burst elect / reconvergence.
    if ((!kPair || crank == 0) && elect_one()) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.n_tile >> 3) << 17) | (((kPair ? 256u : 128u) >> 4) << 24);
      const uint64_t desc_hi = make_smem_desc(0, 128);
      // block windows: the 8-row groups of an A tile are one window row apart
      // A operand: rows of a_sw bytes; block windows: the 8-row groups of an A tile are one window row apart
      const uint64_t desc_hi_a = (make_smem_desc(0, (uint32_t)pp.a_sw) & ~((uint64_t)0x3FFF << 32)) |
                                 ((uint64_t)(((uint32_t)(pp.blk ? p.Wp : 8) * (uint32_t)pp.a_sw) >> 4) << 32);
      const uint32_t a_lo0 = (smem_u32(smem_a) & 0x3FFFFu) >> 4, b_lo0 = (smem_u32(smem_b) & 0x3FFFFu) >> 4;
      const int n_mma = p.taps * (pp.split ? 2 * pp.nh : ((p.Cin_p + 15) >> 4)) * p.T;
      // stacked weights (split): full-width MMAs (N = 2 n_tile, [H | L]) and half-width ones (N = n_tile, into L)
      const uint32_t idesc_n = (1u << 4) | ((uint32_t)(p.n_tile >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t idesc_w = (1u << 4) | ((uint32_t)(p.n_tile >> 2) << 17) | ((128u >> 4) << 24);
      if (pp.b_resident) mbar_wait(w_full, 0);
      uint32_t stage = 0, phase = 0;
      int j = 0;
      for (int w = w0; w < pp.n_windows; w += gridDim.x, ++j) {
        const int slot = j & 1;
        const uint32_t ph = (uint32_t)((j >> 1) & 1);
        const int aslot = pp.a_slots == 2 ? slot : 0;                   // input-window slot (accumulator sets always alternate)
        const uint32_t aph = pp.a_slots == 2 ? ph : (uint32_t)(j & 1);
        if (p.ts && j < 8) p.ts[((size_t)blockIdx.x * 8 + j) * 16 + 0] = gtime();
        mbar_wait(&acc_empty[slot], ph ^ 1u);        // epilogue has drained this accumulator set
        if (p.ts && j < 8) p.ts[((size_t)blockIdx.x * 8 + j) * 16 + 1] = gtime();
        if (!(p.dbg & 8)) mbar_wait(&a_full[aslot], aph);               // window landed
        tc_fence_after();
        if (p.ts && j < 8) {
          p.ts[((size_t)blockIdx.x * 8 + j) * 16 + 2] = gtime();
          p.ts[((size_t)blockIdx.x * 8 + j) * 16 + 6] = (unsigned long long)clock64();
        }
        const uint32_t a_lo = a_lo0 + (((uint32_t)aslot * a_slot) >> 4);
        const uint32_t d0 = tmem_base + (uint32_t)(slot * acc_cols);
        if constexpr (kPair) {
#pragma unroll 4
          for (int i = 0; i < n_mma; ++i) {
            const uint4 e = s_issue[i];
            umma_f16_pair(d0 + e.z, desc_hi_a | (uint64_t)(a_lo + e.x), desc_hi | (uint64_t)(b_lo0 + e.y), idesc, e.w & 1u);
          }
          umma_commit_pair(&a_empty[aslot]);    // both CTAs' window slots may be refilled
          umma_commit_pair(&acc_full[slot]);    // both CTAs' accumulator halves complete
          continue;
        }
        if (pp.chunk_phase) {
          // phase c = the MMAs that read chunk c (chunk 0 already waited for above); each phase releases its chunk
          int i = 0;
          for (int c = 0; c < p.kchunks; ++c) {
            if (c) {
              if (!(p.dbg & 8)) mbar_wait(&a_full[c], aph);
              tc_fence_after();
              if (p.ts && j < 8 && c == 1) p.ts[((size_t)blockIdx.x * 8 + j) * 16 + 9] = gtime();
            }
            const int end = pp.phase_end[c];
#pragma unroll 4
            for (; i < end; ++i) {
              const uint4 e = s_issue[i];
              umma_f16(d0 + e.z, desc_hi_a | (uint64_t)(a_lo + e.x), desc_hi | (uint64_t)(b_lo0 + e.y),
                       pp.split ? ((e.w & 16u) ? idesc_n : idesc_w) : idesc, e.w & 1u);
            }
            umma_commit(&a_empty[c]);
            if (p.ts && j < 8 && c == 0) p.ts[((size_t)blockIdx.x * 8 + j) * 16 + 8] = gtime();
          }
          umma_commit(&acc_full[slot]);
          if (p.ts && j < 8) {
            p.ts[((size_t)blockIdx.x * 8 + j) * 16 + 3] = gtime();
            p.ts[((size_t)blockIdx.x * 8 + j) * 16 + 7] = (unsigned long long)clock64();
          }
          continue;
        }
        if (pp.b_resident) {
#pragma unroll 4
          for (int i = 0; i < n_mma; ++i) {
            const uint4 e = s_issue[i];
            umma_f16(d0 + e.z, desc_hi_a | (uint64_t)(a_lo + e.x), desc_hi | (uint64_t)(b_lo0 + e.y),
                     pp.split ? ((e.w & 16u) ? idesc_n : idesc_w) : idesc, e.w & 1u);
          }
        } else {
          uint32_t b_lo = b_lo0;
          for (int i = 0; i < n_mma; ++i) {
            const uint4 e = s_issue[i];
            if (e.w & 2u) {
              mbar_wait(&b_full[stage], phase);
              tc_fence_after();
              b_lo = b_lo0 + ((stage * b_stage) >> 4);
            }
            umma_f16(d0 + e.z, desc_hi_a | (uint64_t)(a_lo + e.x), desc_hi | (uint64_t)(b_lo + e.y), idesc, e.w & 1u);
            if (e.w & 4u) {
              umma_commit(&b_empty[stage]);
              if (++stage == (uint32_t)p.b_stages) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
        umma_commit(&a_empty[aslot]);     // window slot may be refilled
        umma_commit(&acc_full[slot]);     // accumulators complete
        if (p.ts && j < 8) {
          p.ts[((size_t)blockIdx.x * 8 + j) * 16 + 3] = gtime();
          p.ts[((size_t)blockIdx.x * 8 + j) * 16 + 7] = (unsigned long long)clock64();
        }
      }
    }
  } else if (warp == 3) {
    // ===================== staging DMA (residual in, output out) =====================
    pdl_wait();
    if (pp.n_stage && elect_one()) {
      const int S = pp.n_stage;
      const bool has_res = p.res != nullptr && !(p.dbg & 2);
      const int my_windows = ((int)pp.n_windows - w0 + (int)gridDim.x - 1) / (int)gridDim.x;   // iterations of this CTA (pair)
      auto coords = [&](int jj, int& x0, int& h0, int& b0) {
        const int w = w0 + jj * (int)gridDim.x + w_off;    // past-the-end window of a pair: loads zero-filled, stores clipped
        const int bg = w / p.win_per_img, win = w - bg * p.win_per_img;
        const int wy = pp.blk ? win / pp.wins_x : win;
        x0 = pp.blk ? (win - wy * pp.wins_x) * pp.BW : 0;
        h0 = pp.blk ? wy * pp.BH : win * p.THW;
        b0 = bg * p.TBW;
      };
      // first physical channel of staging block k: hi-plane blocks, then (split storage) the lo-plane blocks
      auto blk_chan = [&](int k) {
        return k < pp.nblk_plane ? n0 + k * pp.cb : p.Cout_p + n0 + (k - pp.nblk_plane) * pp.cb;
      };
      auto fill = [&](int jj) {                 // buffer jj % S is free: fetch window jj's residual (or just release it)
        const int sb = jj % S;
        if (has_res) {
          int x0, h0, b0;
          coords(jj, x0, h0, b0);
          mbar_expect_tx(&stage_full[sb], (uint32_t)pp.nblk * (uint32_t)(pp.rows_stage * pp.cb * 2));
          for (int k = 0; k < pp.nblk; ++k)
            tma_load_4d(smem_stage + (size_t)sb * stage_bytes + (size_t)k * pp.blk_bytes, &map_res, &stage_full[sb],
                        blk_chan(k), x0, h0, b0);
        } else {
          mbar_arrive(&stage_full[sb]);
        }
      };
      for (int jj = 0; jj < S && jj < my_windows; ++jj) fill(jj);
      for (int jj = 0; jj < my_windows; ++jj) {
        const int sb = jj % S;
        mbar_wait(&staged[sb], (uint32_t)((jj / S) & 1));
        if (p.ts && jj < 8) p.ts[((size_t)blockIdx.x * 8 + jj) * 16 + 14] = gtime();
        if (!(p.dbg & 1)) {
          int x0, h0, b0;
          coords(jj, x0, h0, b0);
          for (int k = 0; k < pp.nblk; ++k)
            tma_store_4d(&map_out, smem_stage + (size_t)sb * stage_bytes + (size_t)k * pp.blk_bytes, blk_chan(k), x0, h0,
                         b0);
          bulk_commit();
          bulk_wait_read0();                      // smem of this buffer has been read: it may be refilled
        }
        if (p.ts && jj < 8) p.ts[((size_t)blockIdx.x * 8 + jj) * 16 + 15] = gtime();
        if (jj + S < my_windows) fill(jj + S);
      }
      bulk_wait0();                               // all output writes complete before the CTA exits
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps on every window) =====================
    // Two warps per TMEM lane quarter split each window's (tile, column group) items between them.
    // One epilogue of half the length per window lets the MMA warp alternate accumulator sets without
    // waiting: with one 4-warp group per set the cadence was (t_mma + t_epi) / 2 instead of max(t_mma, t_epi).
    pdl_wait();
    const int half = (warp - 4) >> 2;                  // which of the kParts warps of this lane quarter
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    EpiArgs e{s_bias - n0, p.res, p.out, p.heatmap, p.xs, p.ys, p.coord_maps, p.relu, p.Cout, p.Cout_p, p.H, p.W, p.dbg,
              nullptr, pp.split, (uint32_t)(pp.split ? p.n_tile : p.T * p.n_tile), (uint32_t)(pp.split ? 2 * p.n_tile : 0)};
    const EpiLite el{p.res, p.out, smem_u32(s_bias), p.relu, p.Cout_p, p.dbg, pp.split};
    const uint32_t acc_empty_leader[2] = {kPair ? mapa_rank(&acc_empty[0], 0) : 0u, kPair ? mapa_rank(&acc_empty[1], 0) : 0u};
    int j = 0;
    for (int wb = w0; wb < pp.n_windows; wb += gridDim.x, ++j) {
      const int w = wb + w_off;
      const int slot = j & 1;
      const uint32_t ph = (uint32_t)((j >> 1) & 1);
      const int bg = w / p.win_per_img, win = w - bg * p.win_per_img;
      const int wy = pp.blk ? win / pp.wins_x : win;
      const int x0 = pp.blk ? (win - wy * pp.wins_x) * pp.BW : 0;
      const int h0 = pp.blk ? wy * pp.BH : win * p.THW, b0 = bg * p.TBW;
      if (p.dbg & 64) {
        mbar_wait(&acc_full[slot], ph);
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[slot]);
        continue;
      }
      auto row_of = [&](int t) {
        if (pp.blk) {
          // block windows: lane -> (row & 7, row >> 3) inside tile (t % TX, t / TX); every lane is an output pixel
          EpiRow rr;
          rr.b = b0;
          rr.oh = h0 + 16 * (t / pp.TX) + (row >> 3);
          rr.ow = x0 + 8 * (t % pp.TX) + (row & 7);
          rr.valid = rr.b < p.B && rr.oh < p.H && rr.ow < p.W;
          rr.pix = ((size_t)rr.b * p.H + rr.oh) * p.W + rr.ow;
          return rr;
        }
        // run position -> window position -> (image, row, column); float reciprocals are exact here
        const int pos = t * 128 + row + p.lead;
        const int bi = p.TBW == 1 ? 0 : (int)(((float)pos + 0.5f) * p.inv_img);
        const int rem = pos - bi * p.img_rows;
        const int hp = (int)(((float)rem + 0.5f) * p.inv_wp);
        const int wp = rem - hp * p.Wp;
        EpiRow rr;
        rr.b = b0 + bi;
        rr.oh = h0 + hp - p.halo;
        rr.ow = wp - p.halo;
        rr.valid = pos - p.lead < p.m_run && bi < p.TBW && hp >= p.halo && hp < p.Hw - p.halo && wp >= p.halo &&
                   wp < p.Wp - p.halo && rr.b < p.B && rr.oh < p.H;
        rr.pix = ((size_t)rr.b * p.H + rr.oh) * p.W + rr.ow;
        return rr;
      };
      unsigned long long* ts = (p.ts && j < 8 && half == 0) ? p.ts + ((size_t)blockIdx.x * 8 + j) * 16 : nullptr;
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(slot * acc_cols);
      if constexpr (kHead) {
        e.ts = ts;
        epi_run(e, tbase, p.T, p.n_tile, n0, &acc_full[slot], row_of, ph, half, 2);
      } else {
        if (pp.n_stage) {
          const int sb = j % pp.n_stage;
          const EpiStage es{smem_u32(smem_stage) + (uint32_t)sb * stage_bytes, smem_u32(s_bias), pp.blk_bytes, pp.cb,
                            p.res != nullptr && !(p.dbg & 2), p.relu, p.dbg,
                            pp.split ? (uint32_t)pp.nblk_plane * pp.blk_bytes : 0u};
          mbar_wait(&stage_full[sb], (uint32_t)((j / pp.n_stage) & 1));
          if (ts && (threadIdx.x & 127) == 64) ts[12] = gtime();
          mbar_wait(&acc_full[slot], ph);
          tc_fence_after();
          if (ts && (threadIdx.x & 127) == 64) ts[4] = gtime();
          epi_window_staged<kParts>(es, tbase, p.T, p.n_tile, half, [&](int t) {
            if (pp.blk) {
              // staging buffer = the BW x BH block, row-major; pixels past the image are clipped by the TMA store
              StageRow sr;
              sr.valid = true;
              sr.srow = (uint32_t)((16 * (t / pp.TX) + (row >> 3)) * pp.BW + 8 * (t % pp.TX) + (row & 7));
              return sr;
            }
            const int pos = t * 128 + row + p.lead;
            const int bi = p.TBW == 1 ? 0 : (int)(((float)pos + 0.5f) * p.inv_img);
            const int rem = pos - bi * p.img_rows;
            const int hp = (int)(((float)rem + 0.5f) * p.inv_wp);
            const int wp = rem - hp * p.Wp;
            StageRow sr;
            sr.valid = pos - p.lead < p.m_run && bi < p.TBW && hp >= p.halo && hp < p.Hw - p.halo && wp >= p.halo &&
                       wp < p.Wp - p.halo;
            sr.srow = (uint32_t)(((bi * p.THW) + (hp - p.halo)) * p.W + (wp - p.halo));
            return sr;
          }, pp.ksplit, pp.split);
          if (ts && (threadIdx.x & 127) == 64) ts[13] = gtime();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&staged[sb]);
        } else {
          epi_window_lite<kParts>(el, tbase, p.T, p.n_tile, n0, &acc_full[slot], ph, half, row_of, ts, pp.ksplit, pp.split);
        }
      }
      if (p.ts && j < 8 && lane == 0) atomicMax(p.ts + ((size_t)blockIdx.x * 8 + j) * 16 + 5, gtime());   // last warp done
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                                   // one arrival per warp releases the accumulator set
        if constexpr (kPair) mbar_arrive_cluster(acc_empty_leader[slot]);
        else mbar_arrive(&acc_empty[slot]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();              // the peer's smem / barriers stay valid until both are done
  if (warp == 1) {
    tc_fence_after();
    if constexpr (kPair) tmem_dealloc_pair(tmem_base, p.tmem_cols);
    else tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------
// host side: tensor maps + plan
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct TcConvPlan {
  // shape (batch independent)
  int H, W, Cin_p, OH, OW, Cout_p, Cout, ksize, stride, pad;
  bool split = false;   // fp16x2 storage: cin_a = 2 * Cin_p physical input channels, 2 * Cout_p physical output channels
  int cin_a = 0;        // physical channels of the activation tensor
  int sw;          // swizzle bytes 32 / 64 / 128
  int kc, kchunks, cin_k, n_tile, n_tiles, stages;
  int TW, TH, TB;
  uint32_t tmem_cols;
  size_t smem_bytes;
  // v2 (window run) configuration; use_run == false -> v1 per-tap kernel
  bool use_run = false;
  bool use_persist = false;   // v3: persistent one-CTA-per-SM variant of the window-run kernel
  bool use_tapwin = false;    // v4: tap-window kernel (conv_tapwin_kernel)
  int tw_na = 2, tw_nb = 4;   // its A ring slots / B ring stages
  bool tw_persist = false;    // persistent mode: two accumulator sets, staging buffer of its own (TapWinParams)
  bool tw_pair = false;       // CTA pairs (cta_group::2): each CTA streams half of every weight tile
  int tw_fold = 1;            // N parts folded into the work units (TapWinParams::n_fold)
  uint32_t tw_stage_off = 0, tw_stage_bytes = 0;
  bool use_pair = false;      // v3 with CTA pairs (cta_group::2) instead of an N split over blockIdx.y
  size_t pair_smem = 0;       // dynamic smem per CTA in pair mode
  bool pack_tail = false;     // weight matrix stored with the 32-channel tail chunks of two taps per tile (see PersistParams)
  int w_tiles = 0;            // resident weight tiles per CTA
  int a_slots = 2;            // input-window slots of the persistent kernel (PersistParams::a_slots)
  int b_resident = 0;
  int n_stage = 0, cb = 0, nblk = 0, rows_stage = 0;   // staged (TMA) epilogue of the persistent kernel
  uint32_t blk_bytes = 0;
  int halo = 0, Wp = 0, Hw = 0, THW = 0, TBW = 1, T = 1, rows_alloc = 0, b_stages = 2;
  int blk = 0, BW = 0, BH = 0, TX = 1;                 // block-shaped windows of the persistent kernel (PersistParams)
  int a_sw = 128, a_kchunks = 0;                       // persistent kernel: window row bytes (128 / 64) and chunks per window
  float run_eff = 0.f;
  __half* d_w = nullptr;      // [Cout_p][taps*Cin_p]
  size_t w_bytes = 0;
  CUtensorMap map_b;
  std::map<std::pair<const void*, int>, CUtensorMap> a_maps;  // (input pointer, batch) -> map
  std::map<std::pair<const void*, int>, CUtensorMap> io_maps; // (residual / output pointer, batch) -> staging map
  std::mutex mu;
};

static CUtensorMapSwizzle swizzle_enum(int sw) {
  return sw == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (sw == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

bool tc_conv_supported(const ConvArgs& a) {
  if (a.Cin_p % 16 || a.Cout_p % 16) return false;
  if (a.ksize != 1 && a.ksize != 3) return false;
  if (a.stride != 1 && a.stride != 2) return false;
  if (a.stride == 2 && ((a.H | a.W) & 1)) return false;
  if (a.OW > 128) return false;  // one tile row must fit 128 accumulator lanes
  if (a.Cout_p > 256 && (a.Cout_p % 2 || a.Cout_p / 2 > 256 || (a.Cout_p / 2) % 16)) return false;
  return true;
}

static uint32_t pow2_cols(int n) {
  uint32_t c = 32;
  while ((int)c < n) c <<= 1;
  return c;
}

int tc_conv_plan_create(const ConvArgs& a, const float* wf, TcConvPlan** out) {
  if (!tc_conv_supported(a)) {
    set_error("tc_conv_plan_create: unsupported conv shape");
    return EGN_ERR_INVALID;
  }
  // wf == nullptr: geometry only (kernel choice, tiling, shared-memory / TMEM budget) -- no driver, no device: the
  // CPU test tier pins the plan of every HRNet layer shape through egn_debug_conv_plan
  const bool dry = wf == nullptr;
  EncodeTiledFn enc = dry ? nullptr : get_encode_fn();
  if (!dry && !enc) {
    set_error("cuTensorMapEncodeTiled is not available from the installed driver");
    return EGN_ERR_CUDA;
  }
  TcConvPlan* p = new TcConvPlan();
  p->H = a.H; p->W = a.W; p->Cin_p = a.Cin_p; p->OH = a.OH; p->OW = a.OW;
  p->Cout_p = a.Cout_p; p->Cout = a.Cout; p->ksize = a.ksize; p->stride = a.stride; p->pad = a.pad;
  p->split = a.split != 0;
  const bool split = p->split;
  const int cin_a = split ? 2 * a.Cin_p : a.Cin_p;
  p->cin_a = cin_a;
  // Always 128-byte swizzle rows (64 channels per pipeline stage).  When Cin is not a multiple of
  // 64 the last chunk of a tap is partial: the weight matrix is zero-padded per tap to a multiple
  // of 64 channels, so whatever the activation box holds in those lanes (TMA zero fill past the
  // channel extent, or the neighbouring pixel's channels in the stride-2 view) is multiplied by 0.
  const int force_sw = getenv("EGN_TC_SWIZZLE") ? atoi(getenv("EGN_TC_SWIZZLE")) : 0;
  p->sw = force_sw ? force_sw : 128;
  if (force_sw == 0 && cin_a <= 16) p->sw = 32;
  else if (force_sw == 0 && cin_a <= 32) p->sw = 64;
  p->kc = p->sw / 2;
  p->kchunks = ceil_div(cin_a, p->kc);      // (the fp16x2 per-tap plan recomputes sw / kc / kchunks below)
  // split mode keeps two accumulators per tile (hi*hi | cross terms): at most 256 columns each in v1
  p->n_tiles = a.Cout_p > 256 ? 2 : 1;
  p->n_tile = a.Cout_p / p->n_tiles;
  p->TW = std::min(a.OW, 128);
  p->TH = std::min(a.OH, 128 / p->TW);
  p->TB = p->TH == a.OH ? std::max(1, 128 / (p->TW * p->TH)) : 1;
  p->tmem_cols = pow2_cols((split ? 2 : 1) * p->n_tile);
  const size_t a_stage = 128 * (size_t)p->sw;
  const size_t b_stage = ((size_t)p->n_tile * p->sw + 1023) & ~(size_t)1023;
  const size_t stage = a_stage + b_stage;
  // stage ring budget: small stages -> ~80 KB so that two CTAs share an SM; EGN_TC_V1_BUDGET_KB overrides (tuning)
  // fp16x2 with <= 256 TMEM columns per CTA (96-channel layers: 28 KB stages): three stages and TWO CTAs per SM
  // beat six stages and one (96ch@32x32+res, batch 256: 217 us vs 344 us -- one CTA cannot hide the L2 latency of
  // its own ring); layers that need all 512 columns keep the deep ring (192ch: 183 us vs 217 us).
  const bool two_ctas = split && pow2_cols(2 * p->n_tile) <= 256;
  const size_t budget = getenv("EGN_TC_V1_BUDGET_KB") ? (size_t)atoi(getenv("EGN_TC_V1_BUDGET_KB")) * 1024
                        : (stage > 24 * 1024 ? 176 * 1024 : 80 * 1024);
  const int n_iters = a.ksize * a.ksize * p->kchunks;
  size_t st_count = std::min<size_t>((size_t)kMaxStages, budget / stage);
  st_count = std::min<size_t>(st_count, (size_t)std::max(2, n_iters));
  p->stages = (int)std::max<size_t>(2, st_count);
  p->smem_bytes = 1024 + p->stages * stage + (2 * kMaxStages + 2) * sizeof(uint64_t) + 32 + 4 * (size_t)p->n_tile;
  // ---- v2 window-run configuration (stride-1 convs) ----
  {
    const char* env = getenv("EGN_TC_V2");
    // fp16x2: measured slower than the per-tap kernel (96ch@32x32, batch 256: 545 us vs 345 us -- one CTA per SM and
    // the in-order weight ring expose the L2 latency of every tile); kept behind EGN_TC_V2_SPLIT=1 and tested
    const bool allow = !(env && atoi(env) == 0) && a.stride == 1 && p->sw == 128 && p->kchunks <= kMaxChunks &&
                       (!split || (getenv("EGN_TC_V2_SPLIT") && atoi(getenv("EGN_TC_V2_SPLIT"))));
    const int force_T2 = getenv("EGN_TC_V2_T") ? atoi(getenv("EGN_TC_V2_T")) : 0;
    if (allow) {
      const int halo = a.ksize == 3 ? 1 : 0;
      const int Wp = a.W + 2 * halo, lead = halo * (Wp + 1);
      const int taps_n = a.ksize * a.ksize;
      const size_t b_stage_bytes = ((size_t)p->n_tile * 128 + 1023) & ~(size_t)1023;
      const int n_it = taps_n * p->kchunks;
      float best = 0.f;
      const int nacc2 = split ? 2 : 1;                 // fp16x2: accumulators H and L per M tile
      const int Tmax = std::min(8, 512 / nacc2 / p->n_tile);
      for (int T = 1; T <= Tmax; ++T) {
        if (force_T2 && T != force_T2) continue;
        for (int multi = 0; multi < 2; ++multi) {
          int THW, TBW;
          if (!multi) {
            THW = std::min(a.H, (128 * T + 2 * halo) / Wp);
            TBW = 1;
            if (THW < 1) continue;
          } else {
            THW = a.H;
            const int img = (a.H + 2 * halo) * Wp;
            TBW = (128 * T + 2 * lead) / img;
            if (TBW < 2) continue;
            TBW = std::min(TBW, 64);
          }
          const int Hw = THW + 2 * halo;
          if (Wp > 256 || Hw > 256) continue;
          const int rows_win = TBW * Hw * Wp;
          const int m_run = rows_win - 2 * lead;
          if (m_run > 128 * T) continue;
          const int rows_alloc = (std::max(T * 128 + 2 * lead, rows_win) + 7) & ~7;
          const size_t a_bytes = (size_t)p->kchunks * rows_alloc * 128;
          int bst = std::min(kMaxBStages, std::max(2, n_it));
          size_t smem = 1024 + a_bytes + bst * b_stage_bytes + 1536;
          while (smem > 200 * 1024 && bst > 2) {
            --bst;
            smem = 1024 + a_bytes + bst * b_stage_bytes + 1536;
          }
          if (smem > 200 * 1024) continue;
          // cost model (nominal batch 64): CTAs per SM x (MMA rows + window load + fixed overhead);
          // two co-resident CTAs overlap each other's load / epilogue phases
          const int windows = multi ? 1 : ceil_div(a.H, THW);
          const double eff = multi ? (double)TBW * a.H * a.W / ((double)T * 128)
                                   : (double)a.H * a.W / ((double)windows * T * 128);
          const bool two_per_sm = smem <= 112 * 1024 && pow2_cols(nacc2 * T * p->n_tile) <= 256;
          const int ctas = windows * ceil_div(64, TBW) * p->n_tiles;
          const double cta_cost = (split ? 1.5 : 1.0) * T + 0.6 * rows_win / 128.0 * (split ? 2 : 1) + 0.7;
          const double est = ceil_div(ctas, 148) * cta_cost * (two_per_sm ? 0.7 : 1.0);
          const float score = (float)(1.0 / est);
          if (score > best) {
            best = score;
            p->use_run = true;
            p->halo = halo; p->Wp = Wp; p->Hw = Hw; p->THW = THW; p->TBW = TBW; p->T = T;
            p->rows_alloc = rows_alloc; p->b_stages = bst; p->run_eff = (float)eff;
            p->smem_bytes = smem;
            p->tmem_cols = pow2_cols(nacc2 * T * p->n_tile);
          }
        }
      }
      // v1 keeps the small, weight-dominated maps where the run's junk rows cost more than they save
      const char* minw = getenv("EGN_TC_V2_MIN_W");
      if (p->use_run && a.W < (minw ? atoi(minw) : 24) && a.ksize == 3) p->use_run = false;
      if (p->use_run && p->run_eff < 0.6f) p->use_run = false;
      if (getenv("EGN_TC_VERBOSE"))
        fprintf(stderr, "[egn] conv %dx%d s%d %d->%d @%dx%d: %s T=%d THW=%d TBW=%d eff=%.2f smem=%zuKB bst=%d tmem=%u\n", a.ksize, a.ksize,
                a.stride, a.Cin_p, a.Cout_p, a.H, a.W, p->use_run ? "v2-run" : "v1-tap", p->T, p->THW, p->TBW, p->run_eff,
                p->smem_bytes / 1024, p->b_stages, p->tmem_cols);
      if (!p->use_run) {
        p->tmem_cols = pow2_cols((split ? 2 : 1) * p->n_tile);
        p->smem_bytes = 1024 + p->stages * stage + (2 * kMaxStages + 2) * sizeof(uint64_t) + 32 + 4 * (size_t)p->n_tile;
      }
    }
  }
  // ---- v3 persistent configuration: double-buffered windows + accumulators, weights RESIDENT in smem ----
  // If the whole [Cout][taps*Cin] matrix does not fit next to two window slots, the output channels are
  // split over blockIdx.y (each half keeps its own half of the weights resident; the window is loaded by
  // both).  Streaming the weights per window instead costs more L2->SM bandwidth than it saves.
  {
    const char* env = getenv("EGN_TC_V3");
    const bool allow = !(env && atoi(env) == 0) && a.stride == 1 && p->sw == 128 && p->kchunks <= kMaxChunksPersist &&
                       (a.ksize == 1 ? a.W >= 32 : a.W >= 24);
    const int nacc = split ? 2 : 1;                  // accumulators per M tile
    const int kslices = (cin_a + 15) / 16;           // K16 slices of a tap's weight row
    const int groups = split ? 2 * (a.Cin_p / 16) : kslices;   // MMA groups per tap (split: stacked weights, PersistParams)
    if (allow) {
      const int halo = a.ksize == 3 ? 1 : 0;
      const int Wp = a.W + 2 * halo, lead = halo * (Wp + 1);
      // split: [w_hi | w_lo] row-stacked tiles over densely packed K16 slices (four per tile)
      const int w_tiles = split ? ceil_div(a.ksize * a.ksize * (a.Cin_p / 16), 4) : a.ksize * a.ksize * p->kchunks;
      const size_t smem_cap = 225 * 1024;
      double best = 1e30;
      const int force_T = getenv("EGN_TC_V3_T") ? atoi(getenv("EGN_TC_V3_T")) : 0;
      const int max_stage = getenv("EGN_TC_STAGE") ? atoi(getenv("EGN_TC_STAGE")) : 2;
      const int base_tiles = p->n_tiles;
      // Splitting the output channels over blockIdx.y (split = 2) lets each half keep its weights resident
      // when the whole matrix does not fit (96 channels: 2 x 108 KB); both halves then load the window.
      // (fp16x2, EGN_TC_V3_NSPLIT=3: 96 channels as 3 x 32 keep 112 KB of stacked weights each -- measured 186 us
      // against 156 us for the tap-window kernel with full-width N = 192 / 96 MMAs: off by default)
      const int force_nsplit = getenv("EGN_TC_V3_NSPLIT") ? atoi(getenv("EGN_TC_V3_NSPLIT")) : 0;
      for (int nsplit = 1; nsplit <= (getenv("EGN_TC_V3_NOSPLIT") ? 1 : (force_nsplit ? 4 : 2)); ++nsplit) {
        if (force_nsplit && nsplit != force_nsplit) continue;
        if (!split && nsplit == 3) continue;
        const int n_tiles = base_tiles * nsplit;
        if (a.Cout_p % (16 * n_tiles)) continue;
        const int n_tile = a.Cout_p / n_tiles;
        if (n_tile > 256 / nacc || n_tile < 16) continue;
        const size_t b_stage_bytes = ((size_t)n_tile * (split ? 256 : 128) + 1023) & ~(size_t)1023;
        const int Tmax = std::min(8, 256 / nacc / n_tile);
        // window shapes: full-width rows of one image (mode 0), whole images (mode 1), or -- 3x3 convs -- blocks of
        // 8 * TX x 16 * TY pixels whose M tiles are 8 columns x 16 rows (mode 2 + TY: PersistParams::blk; every
        // accumulator row is an output pixel)
        const bool allow_blk = a.ksize == 3 && !(getenv("EGN_TC_BLK") && atoi(getenv("EGN_TC_BLK")) == 0);
        for (int T = 1; T <= Tmax; ++T) {
          for (int mode = 0; mode < 2 + (allow_blk ? T : 0); ++mode) {
            const int multi = mode == 1;
            const int TY = mode >= 2 ? mode - 1 : 0;          // block mode: tiles down (TX = T / TY across)
            if (TY && T % TY) continue;
            const int blk = TY ? 1 : 0, TX = TY ? T / TY : 1, BW = 8 * TX, BH = 16 * TY;
            if (blk && (BW >= a.W + 8 || BH >= a.H + 16)) continue;          // a block larger than the map
            if (getenv("EGN_TC_BLK") && atoi(getenv("EGN_TC_BLK")) == 2 && allow_blk && !blk) continue;   // tuning: blocks only
            int THW, TBW;
            if (blk) {
              THW = BH;
              TBW = 1;
            } else if (!multi) {
              THW = std::min(a.H, (128 * T + 2 * halo) / Wp);
              TBW = 1;
              if (THW < 1) continue;
            } else {
              THW = a.H;
              TBW = (128 * T + 2 * lead) / ((a.H + 2 * halo) * Wp);
              if (TBW < 2) continue;
              TBW = std::min(TBW, 64);
            }
            const int Hw = THW + 2 * halo;
            const int Wpw = blk ? BW + 2 * halo : Wp;          // window pitch in pixels
            if (Wpw > 256 || Hw > 256) continue;
            const int rows_win = TBW * Hw * Wpw;
            if (!blk && rows_win - 2 * lead > 128 * T) continue;
            const int rows_alloc = blk ? (rows_win + 15) & ~15 : (std::max(T * 128 + 2 * lead, rows_win) + 15) & ~15;
            // fp16x2 windows are twice the bytes: also try a single window slot (the measured fp16 plans keep two),
            // and 64-byte window rows when the physical channels are an odd number of 32-channel chunks (96 = 3 x 32)
            const int force_slots = getenv("EGN_TC_ASLOTS") ? atoi(getenv("EGN_TC_ASLOTS")) : 0;
            const int force_asw = getenv("EGN_TC_ASW") ? atoi(getenv("EGN_TC_ASW")) : 0;
            for (int a_sw = 128; a_sw >= (split && cin_a % 64 == 32 ? 64 : 128); a_sw -= 64) {
            if (force_asw && a_sw != force_asw && (force_asw == 128 || (split && cin_a % 64 == 32))) continue;
            const int a_kch = ceil_div(cin_a, a_sw / 2);
            if (a_kch > kMaxChunksPersist) continue;
            for (int a_slots = 2; a_slots >= (split ? 1 : 2); --a_slots) {
            if (force_slots && a_slots != force_slots) continue;
            if (a_slots == 1 && (a_kch < 2 || a_kch > 4)) continue;       // one slot = a chunk-phased ring of 2..4 chunks
            const size_t a_bytes = (size_t)a_slots * a_kch * rows_alloc * a_sw;
            const size_t fixed = 1024 + 320 + (size_t)n_tile * 4 +
                                 (size_t)a.ksize * a.ksize * groups * T * 16;   // barriers, bias, issue table
            // resident weights: the 32-channel tail chunks of a 3x3 conv are packed two taps per tile
            const bool can_pack = !split && a.ksize == 3 && p->kchunks >= 2 && cin_a % 64 == 32 && !getenv("EGN_TC_NOPACK");
            const int w_tiles_res = can_pack ? a.ksize * a.ksize * (p->kchunks - 1) + (a.ksize * a.ksize + 1) / 2 : w_tiles;
            size_t smem = a_bytes + (size_t)w_tiles_res * b_stage_bytes + fixed;
            int resident = 1, bst = 0;
            if (smem > smem_cap) {
              // weights streamed through a ring, once per window: only worth it when a window holds >= 2 M tiles
              // (fp16x2: 2x the tiles per window -- measured 512 us vs 345 us for the per-tap kernel on 96ch@32x32)
              if (T < 2 || split || getenv("EGN_TC_V3_NOSTREAM")) continue;
              resident = 0;
              bst = std::min(kMaxBStages, w_tiles);
              smem = a_bytes + bst * b_stage_bytes + fixed;
              while (smem > smem_cap && bst > 3) {
                --bst;
                smem = a_bytes + bst * b_stage_bytes + fixed;
              }
              if (smem > smem_cap) continue;
            }
            const int windows = blk ? ceil_div(a.W, BW) * ceil_div(a.H, BH) : (multi ? 1 : ceil_div(a.H, THW));
            const double eff = multi ? (double)TBW * a.H * a.W / ((double)T * 128)
                                     : (double)a.H * a.W / ((double)windows * T * 128);
            if (force_T && T != force_T) continue;
            // staging buffers of the TMA epilogue: [block][TBW*THW*W pixels][cb channels]
            const int cb = n_tile % 64 == 0 ? 64 : (n_tile % 48 == 0 ? 48 : (n_tile % 32 == 0 ? 32 : 0));
            const int rows_stage = blk ? BW * BH : TBW * THW * a.W;
            const size_t blk_bytes = cb ? (((size_t)rows_stage * cb * 2 + 1023) & ~(size_t)1023) : 0;
            const int nblk = cb ? n_tile / cb : 0;
            // cost model in SM cycles at the nominal batch 256 (measured constants, profiles/r01_*):
            //   tcgen05.mma 128 x N x 16: 40 cycles up to N=32, 44 @48, 48 @64, 56 @96, 64 @128, 118 @192, 150 @256;
            //     x1.5 while a direct (uncoalesced) epilogue competes for the smem / L1 data path
            //   epilogue: ~250 (direct) / ~100 (staged) cycles per 16-column item and lane quarter
            //   a single staging buffer exposes store drain + residual load (~4000 cycles measured) once per window
            //   streamed weights: ~2400 cycles per ring tile (measured on 48- and 96-channel layers; the ring does
            //   not hide the L2 latency behind the window loads queued on the same TMA unit) -> last resort
            const int n_win = windows * ceil_div(256, TBW);
            const double mma_cyc = n_tile <= 32 ? 40.0 : n_tile <= 64 ? 40.0 + (n_tile - 32) * 0.25
                                 : n_tile <= 128 ? 48.0 + (n_tile - 64) * 0.25 : 64.0 + (n_tile - 128) * 0.68;
            // split: per slice one 2 n_tile-wide and one n_tile-wide MMA (operand-read bound: 32 + N / 4 cycles)
            const double t_mma = split ? (double)a.ksize * a.ksize * (a.Cin_p / 16) * T * (64.0 + 0.75 * n_tile)
                                       : (double)a.ksize * a.ksize * groups * T * mma_cyc;
            const double t_w = resident ? 0.0 : (double)w_tiles * 2400.0;
            const double t_a = (double)a_kch * rows_win * a_sw / 48.0;      // window load, ~48 B/cycle/SM
            const int min_stage = getenv("EGN_TC_STAGE_MIN") ? atoi(getenv("EGN_TC_STAGE_MIN")) : 0;   // tuning
            for (int S = max_stage; S >= min_stage; --S) {
              if (S && !cb) continue;
              const size_t smem_s = smem + (size_t)S * nblk * blk_bytes * nacc;     // split: hi and lo planes staged
              if (smem_s > smem_cap) continue;
              const double t_epi = (double)T * (n_tile / 16) * (S ? 100.0 : 250.0) * nacc;
              const double t_mm = (S ? t_mma : 1.5 * t_mma) + (a_slots == 1 ? t_a / a_kch : 0.0);   // one slot: a chunk ring
              const double t_win = std::max(std::max(t_mm, t_epi), std::max(t_w, t_a)) +
                                   (S == 1 ? 4000.0 : 0.0) + 300.0;
              const double est = ceil_div(n_win * n_tiles, 148) * t_win;
              if (est < best) {
                best = est;
                p->use_persist = true;
                p->b_resident = resident;
                p->n_tiles = n_tiles;
                p->n_tile = n_tile;
                p->halo = halo; p->Wp = Wpw; p->Hw = Hw; p->THW = THW; p->TBW = TBW; p->T = T;
                p->blk = blk; p->BW = BW; p->BH = BH; p->TX = TX;
                p->rows_alloc = rows_alloc; p->b_stages = bst; p->run_eff = (float)eff;
                p->smem_bytes = smem_s;
                p->tmem_cols = pow2_cols(2 * nacc * T * n_tile);
                p->n_stage = S; p->cb = cb; p->nblk = nblk; p->rows_stage = rows_stage; p->blk_bytes = (uint32_t)blk_bytes;
                p->pack_tail = resident && can_pack;
                p->w_tiles = resident ? w_tiles_res : w_tiles;
                p->a_slots = a_slots;
                p->a_sw = a_sw;
                p->a_kchunks = a_kch;
              }
            }
            }
            }
          }
        }
      }
      if (p->use_persist && p->run_eff < 0.6f) {
        p->use_persist = false;
        if (p->blk) p->use_run = false;          // the window-run plan's geometry was overwritten by the block candidate
        p->blk = 0;
        p->pack_tail = false;
        p->n_tiles = base_tiles;
        p->n_tile = a.Cout_p / base_tiles;
      }
      if (p->use_persist) p->use_run = false;
      // An N split chosen only to make the weights fit (two passes of N/2-wide MMAs over the same windows) runs
      // as ONE pass of CTA pairs instead: same smem per CTA, N-wide cta_group::2 MMAs (EGN_TC_PAIR=0 disables).
      {
        const char* pe = getenv("EGN_TC_PAIR");
        // per CTA the pair keeps the split's weights and windows but stages / biases the full N
        const size_t pair_smem = p->smem_bytes + (size_t)p->n_stage * p->nblk * p->blk_bytes * nacc + 4 * (size_t)p->n_tile;
        p->use_pair = p->use_persist && !split && !(pe && atoi(pe) == 0) && base_tiles == 1 && p->n_tiles == 2 && p->b_resident &&
                      (2 * p->n_tile) % 16 == 0 && 2 * nacc * p->T * 2 * p->n_tile <= 512 && pair_smem <= 227 * 1024;
        if (p->use_pair) p->pair_smem = pair_smem;
        if (p->use_pair) p->tmem_cols = pow2_cols(2 * nacc * p->T * 2 * p->n_tile);
      }
      if (getenv("EGN_TC_VERBOSE") && p->use_persist && p->use_pair) fprintf(stderr, "[egn] (next line) CTA-pair mode\n");
      if (getenv("EGN_TC_VERBOSE") && p->use_persist)
        fprintf(stderr, "[egn] conv %dx%d s%d %d->%d @%dx%d%s: v3-persist blk=%dx%d a_sw=%d a_slots=%d T=%d THW=%d TBW=%d eff=%.2f smem=%zuKB n_tiles=%d n_tile=%d resident=%d bst=%d tmem=%u stage=%dx%dx%uB cb=%d\n",
                a.ksize, a.ksize, a.stride, a.Cin_p, a.Cout_p, a.H, a.W, split ? " fp16x2" : "", p->blk ? p->BW : 0, p->blk ? p->BH : 0, p->a_sw, p->a_slots, p->T, p->THW, p->TBW, p->run_eff,
                p->smem_bytes / 1024, p->n_tiles, p->n_tile, p->b_resident, p->b_stages, p->tmem_cols, p->n_stage, p->nblk,
                p->blk_bytes, p->cb);
    }
  }
  if (!p->use_persist && !p->use_run) {
    // v1 after a rejected window-run / persistent plan: its own TMEM and shared-memory footprint
    // (A single accumulator for N = 192 tiles -- cross terms added to the hi*hi sum so that two CTAs fit per SM -- was
    // measured at +2 % end to end but 2.5x the error, past the 1e-4 bound on 3-D key-points: removed.)
    p->tmem_cols = pow2_cols((split ? 2 : 1) * p->n_tile);
    // v4 tap-window kernel: stride-1 3x3 convs whose map tiles into 8 x 16 blocks without much waste
    {
      const char* e4 = getenv("EGN_TC_V4");
      const int tiles = ceil_div(a.W, 8) * ceil_div(a.H, 16);
      const double eff4 = (double)a.H * a.W / ((double)tiles * 128);
      const int v4_mode = e4 ? atoi(e4) : 1;              // 0 off, 1 fp16x2 only, 2 plain fp16 too
      if (a.ksize == 3 && a.stride == 1 && a.pad == 1 && eff4 >= 0.75 && v4_mode && (split || v4_mode == 2) && force_sw == 0) {
        p->use_tapwin = true;
        // N fold: 129..256 output channels run as two parts of <= 128 so that two accumulator sets fit in TMEM and
        // the tile loop can be persistent (EGN_TC_V4_FOLD=0 disables)
        if (p->n_tile > 128 && p->n_tile <= 256 && (p->n_tile / 2) % 32 == 0 &&
            !(getenv("EGN_TC_V4_FOLD") && atoi(getenv("EGN_TC_V4_FOLD")) == 0)) {
          p->tw_fold = 2;
          p->n_tile /= 2;
          p->tmem_cols = pow2_cols((split ? 2 : 1) * p->n_tile);
        }
        // CTA pairs: half of the weight rows per CTA (EGN_TC_V4_PAIR=0 disables)
        p->tw_pair = p->n_tile % 32 == 0 && !(getenv("EGN_TC_V4_PAIR") && atoi(getenv("EGN_TC_V4_PAIR")) == 0);
        // 64-channel chunks, except 32-channel ones for single CTAs of <= 256 TMEM columns (two per SM or persistent with
        // 12 KB stages; measured, batch 256: 96ch 156 vs 226 us unpaired, 192ch 138 vs 128 us).  The kernel is bound by
        // the bytes its weight ring keeps in flight against the ~1.2 us L2 -> smem latency under load: paired 96ch
        // 118 / 114 / 113 us with 8 / 12 / 16 stages of 6 KB, 110 us with 6 stages of 12 KB.
        p->sw = getenv("EGN_TC_V4_SW") ? atoi(getenv("EGN_TC_V4_SW")) : ((p->tmem_cols <= 256 && !p->tw_pair) ? 64 : 128);
        p->kc = p->sw / 2;
        p->kchunks = ceil_div(a.Cin_p, p->kc);
        const size_t a_plane = ((size_t)kTwWp * kTwHp * p->sw + 1023) & ~(size_t)1023;
        const size_t a_slot = (split ? 2 : 1) * a_plane;
        const size_t b_stage4 = ((size_t)(p->tw_pair ? p->n_tile / 2 : p->n_tile) * p->sw * (split ? 2 : 1) + 1023) & ~(size_t)1023;
        const bool two = p->tmem_cols <= 256;
        // staged epilogue: [block][128 pixels][cb channels], hi blocks then lo blocks
        p->cb = p->n_tile % 64 == 0 ? 64 : (p->n_tile % 48 == 0 ? 48 : (p->n_tile % 32 == 0 ? 32 : 0));
        p->nblk = p->cb ? p->n_tile / p->cb : 0;
        p->blk_bytes = (uint32_t)(128 * p->cb * 2);
        const size_t stage_bytes = (size_t)(split ? 2 : 1) * p->nblk * p->blk_bytes;
        const bool want_staged = p->cb && !(getenv("EGN_TC_V4_STAGED") && atoi(getenv("EGN_TC_V4_STAGED")) == 0);
        // persistent mode: two accumulator sets (<= 256 columns each) and the staging buffer behind the rings
        p->tw_persist = two && want_staged && !(getenv("EGN_TC_V4_PERSIST") && atoi(getenv("EGN_TC_V4_PERSIST")) == 0);
        const size_t fixed4 = 1024 + (2 * kTwMaxA + 2 * kTwMaxB + 6) * sizeof(uint64_t) + 32 + (size_t)p->n_tile * p->tw_fold * 4;
        const size_t budget4 = getenv("EGN_TC_V4_BUDGET_KB") ? (size_t)atoi(getenv("EGN_TC_V4_BUDGET_KB")) * 1024
                                                             : (p->tw_persist ? 225 * 1024 - stage_bytes : (two ? 106 * 1024 : 224 * 1024));
        p->tw_na = getenv("EGN_TC_V4_NA") ? atoi(getenv("EGN_TC_V4_NA")) : 2;      // (3 measured equal once the refill moved to tap 3)
        // at most one slot per chunk of a tile -- but a persistent CTA walks over many tiles: never fewer than two slots
        // there (a one-chunk layer would otherwise refill the slot it is still multiplying from)
        p->tw_na = std::max(1, std::min(std::min(p->tw_na, kTwMaxA), std::max(p->tw_persist ? 2 : 1, p->kchunks)));
        if (p->tw_persist) p->tw_na = std::max(2, p->tw_na);
        long nb = ((long)budget4 - (long)fixed4 - (long)(p->tw_na * a_slot)) / (long)b_stage4;
        if (getenv("EGN_TC_V4_NB")) nb = std::min<long>(nb, atoi(getenv("EGN_TC_V4_NB")));
        p->tw_nb = (int)std::max<long>(2, std::min<long>(nb, kTwMaxB));
        const size_t rings = p->tw_na * a_slot + p->tw_nb * b_stage4;
        p->tw_stage_off = p->tw_persist ? (uint32_t)rings : 0u;
        p->tw_stage_bytes = p->tw_persist ? (uint32_t)stage_bytes : 0u;
        p->smem_bytes = fixed4 + rings + p->tw_stage_bytes;
        if (p->smem_bytes > 227 * 1024) {
          // does not fit: back to the per-tap kernel with the full tile width
          p->use_tapwin = false;
          p->n_tile *= p->tw_fold;
          p->tw_fold = 1;
          p->tw_pair = p->tw_persist = false;
          p->tmem_cols = pow2_cols((split ? 2 : 1) * p->n_tile);
        }
        p->n_stage = (p->use_tapwin && want_staged && (p->tw_persist || stage_bytes <= rings)) ? 1 : 0;
        if (p->tw_persist) p->tmem_cols = pow2_cols(2 * (split ? 2 : 1) * p->n_tile);
        if (p->use_tapwin) {
          p->blk = 1; p->BW = 8; p->BH = 16; p->TBW = 1;        // staging box of make_io_map
        }
        if (getenv("EGN_TC_VERBOSE") && p->use_tapwin)
          fprintf(stderr, "[egn] conv %dx%d s%d %d->%d @%dx%d%s: v4-tapwin n_tile=%d sw=%d kchunks=%d na=%d nb=%d smem=%zuKB tmem=%u eff=%.2f staged=%d cb=%d persist=%d pair=%d fold=%d\n",
                  a.ksize, a.ksize, a.stride, a.Cin_p, a.Cout_p, a.H, a.W, split ? " fp16x2" : "", p->n_tile, p->sw, p->kchunks,
                  p->tw_na, p->tw_nb, p->smem_bytes / 1024, p->tmem_cols, eff4, p->n_stage, p->cb, p->tw_persist ? 1 : 0, p->tw_pair ? 1 : 0, p->tw_fold);
      }
    }
    if (p->use_tapwin) {
      // (geometry fixed above)
    } else if (split) {
      // fat stages (TcParams): x_hi box + x_lo box + [w_hi | w_lo] rows per (tap, kc LOGICAL channels).  32-channel
      // stages by default (EGN_TC_V1_SW overrides): more, smaller stages in flight for the same bytes.  Tiles of at most
      // 256 TMEM columns: ~100 KB rings so that TWO CTAs share an SM (one CTA cannot hide the L2 latency of its own ring);
      // wider tiles need all 512 columns: one CTA with the deepest ring that fits.
      if (force_sw == 0) p->sw = getenv("EGN_TC_V1_SW") ? atoi(getenv("EGN_TC_V1_SW")) : 64;
      p->kc = p->sw / 2;
      p->kchunks = ceil_div(a.Cin_p, p->kc);
      const size_t stage2 = 2 * 128 * (size_t)p->sw + (((size_t)2 * p->n_tile * p->sw + 1023) & ~(size_t)1023);
      const size_t budget2 = getenv("EGN_TC_V1_BUDGET_KB") ? (size_t)atoi(getenv("EGN_TC_V1_BUDGET_KB")) * 1024
                                                           : (two_ctas ? 104 * 1024 : 208 * 1024);
      size_t st2 = std::min<size_t>((size_t)kMaxStages, budget2 / stage2);
      st2 = std::min<size_t>(st2, (size_t)std::max(2, a.ksize * a.ksize * p->kchunks));
      p->stages = (int)std::max<size_t>(2, st2);
      p->smem_bytes = 1024 + p->stages * stage2 + (2 * kMaxStages + 2) * sizeof(uint64_t) + 32 + 4 * (size_t)p->n_tile;
    } else {
      p->smem_bytes = 1024 + p->stages * stage + (2 * kMaxStages + 2) * sizeof(uint64_t) + 32 + 4 * (size_t)p->n_tile;
    }
  }
  if (!p->use_persist && !p->use_run && !p->use_tapwin) {
    // staged TMA epilogue over the dead stage ring: [block][TW*TH*TB pixels][cb channels], hi blocks then lo blocks
    const size_t a_st = 128 * (size_t)p->sw * (split ? 2 : 1);
    const size_t b_st = ((size_t)p->n_tile * p->sw * (split ? 2 : 1) + 1023) & ~(size_t)1023;
    p->cb = p->n_tile % 64 == 0 ? 64 : (p->n_tile % 48 == 0 ? 48 : (p->n_tile % 32 == 0 ? 32 : (p->n_tile % 16 == 0 ? 16 : 0)));
    p->nblk = p->cb ? p->n_tile / p->cb : 0;
    p->blk_bytes = (uint32_t)(((size_t)p->TW * p->TH * p->TB * p->cb * 2 + 1023) & ~(size_t)1023);
    p->n_stage = (p->cb && !(getenv("EGN_TC_V1_STAGED") && atoi(getenv("EGN_TC_V1_STAGED")) == 0) &&
                  (size_t)(split ? 2 : 1) * p->nblk * p->blk_bytes <= p->stages * (a_st + b_st)) ? 1 : 0;
  }
  if (getenv("EGN_TC_VERBOSE") && !p->use_persist && !p->use_run && !p->use_tapwin)
    fprintf(stderr, "[egn] conv %dx%d s%d %d->%d @%dx%d%s: v1-tap n_tile=%d sw=%d stages=%d smem=%zuKB tmem=%u staged=%d cb=%d\n", a.ksize, a.ksize,
            a.stride, a.Cin_p, a.Cout_p, a.H, a.W, split ? " fp16x2" : "", p->n_tile, p->sw, p->stages, p->smem_bytes / 1024, p->tmem_cols,
            p->n_stage, p->cb);
  // weights: folded [tap][Cin_p][Cout_p] fp32 -> fp16 matrices, w_hi = rn16(w), w_lo = rn16(w - w_hi):
  //   plain fp16                 [Cout_p][taps * cin_k], tap row = [w (Cin_p)] zero padded to whole kc-channel chunks
  //                              (persistent kernel: 32-channel tail chunks of two taps packed per tile, pack_tail)
  //   fp16x2, v1 / v3 / v4       the STACKED matrix [2 * Cout_p][taps * Cin_p rounded up to 64]: rows [0, Cout_p) = w_hi,
  //                              [Cout_p, 2 Cout_p) = w_lo, K index tap * Cin_p + c (K16 slices packed densely, no per-tap
  //                              padding; PersistParams / TcParams / TapWinParams)
  //   fp16x2, v2 (window-run)    [Cout_p][taps * cin_k], tap row = [w_hi (Cin_p) | w_lo (Cin_p)] (cross pairing by the issuer)
  const int taps = a.ksize * a.ksize;
  const int cin_k = p->kchunks * p->kc;
  p->cin_k = cin_k;
  const bool pack = p->use_persist && p->pack_tail;
  const bool v3_split = split && !p->use_run;                           // the stacked matrix
  const int full_k = (p->kchunks - 1) * 64;              // channels of a tap that live in full 64-wide chunks
  const size_t tap_k = (size_t)cin_k;
  const size_t K = v3_split ? ((size_t)taps * a.Cin_p + 63) / 64 * 64
                            : (pack ? (size_t)taps * full_k + (size_t)((taps + 1) / 2) * 64 : (size_t)taps * tap_k);
  const size_t w_rows = v3_split ? 2 * (size_t)a.Cout_p : (size_t)a.Cout_p;
  if (dry) {
    p->w_bytes = w_rows * K * sizeof(__half);
    *out = p;
    return EGN_OK;
  }
  std::vector<__half> w(w_rows * K, __float2half_rn(0.f));
  for (int o = 0; o < a.Cout_p; ++o)
    for (int t = 0; t < taps; ++t)
      for (int c = 0; c < a.Cin_p; ++c) {
        const float wv = wf[((size_t)t * a.Cin_p + c) * a.Cout_p + o];
        const __half hi = __float2half_rn(wv);
        const __half lo = __float2half_rn(wv - __half2float(hi));
        auto put = [&](int pos, __half v) {              // pos: channel position inside the tap row
          size_t k = (size_t)t * tap_k + pos;
          if (pack) k = pos < full_k ? (size_t)t * full_k + pos
                                     : (size_t)taps * full_k + (size_t)(t >> 1) * 64 + (size_t)(t & 1) * 32 + (pos - full_k);
          w[(size_t)o * K + k] = v;
        };
        if (v3_split) {
          w[(size_t)o * K + (size_t)t * a.Cin_p + c] = hi;
          w[((size_t)a.Cout_p + o) * K + (size_t)t * a.Cin_p + c] = lo;
        } else if (!split) {
          put(c, hi);
        } else {
          put(c, hi);
          put(a.Cin_p + c, lo);
        }
      }
  p->w_bytes = w.size() * sizeof(__half);
  if (cudaMalloc(&p->d_w, p->w_bytes) != cudaSuccess ||
      cudaMemcpy(p->d_w, w.data(), p->w_bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("tc_conv_plan_create: weight upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    tc_conv_plan_destroy(p);
    return EGN_ERR_CUDA;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)w_rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {(cuuint32_t)p->kc, (cuuint32_t)(p->use_tapwin && p->tw_pair ? p->n_tile / 2 : p->n_tile)};
  const cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&p->map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, p->d_w, gdim, gstr, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(p->sw), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
    tc_conv_plan_destroy(p);
    return EGN_ERR_CUDA;
  }
  *out = p;
  return EGN_OK;
}

void tc_conv_plan_destroy(TcConvPlan* p) {
  if (!p) return;
  cudaFree(p->d_w);
  delete p;
}

size_t tc_conv_plan_weight_bytes(const TcConvPlan* p) { return p ? p->w_bytes : 0; }

// [B][OH][OW][Cout_p] tensor seen through the staging box of the persistent kernel's TMA epilogue
static int make_io_map(TcConvPlan* p, const void* ptr, int B, CUtensorMap* m) {
  EncodeTiledFn enc = get_encode_fn();
  const cuuint64_t C = (p->split ? 2 : 1) * p->Cout_p, W = p->OW, H = p->OH;     // fp16x2: [hi | lo] planes
  const cuuint64_t gdim[4] = {C, W, H, (cuuint64_t)B};
  const cuuint64_t gstr[3] = {C * 2, W * C * 2, H * W * C * 2};
  const bool v1 = !p->use_persist && !p->use_run && !p->use_tapwin;       // per-tap kernel: the output tile
  const cuuint32_t box[4] = {(cuuint32_t)p->cb, (cuuint32_t)(v1 ? p->TW : (p->blk ? p->BW : p->OW)),
                             (cuuint32_t)(v1 ? p->TH : (p->blk ? p->BH : p->THW)), (cuuint32_t)(v1 ? p->TB : p->TBW)};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, p->cb == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(staging %dx%dx%d, B=%d, box %dx%dx%dx%d) failed with %d", p->OH, p->OW, p->Cout_p, B,
              p->cb, p->OW, p->THW, p->TBW, (int)r);
    return EGN_ERR_CUDA;
  }
  return EGN_OK;
}

static int make_a_map(TcConvPlan* p, const void* in, int B, CUtensorMap* m) {
  EncodeTiledFn enc = get_encode_fn();
  const cuuint64_t C = p->cin_a, W = p->W, H = p->H;      // physical channels (fp16x2: [hi | lo] planes)
  CUresult r;
  if (p->use_run || p->use_persist) {
    const int a_sw = p->use_persist ? p->a_sw : 128;          // window row bytes: 64 channels (SW128) or 32 (SW64)
    const cuuint64_t gdim[4] = {C, W, H, (cuuint64_t)B};
    const cuuint64_t gstr[3] = {C * 2, W * C * 2, H * W * C * 2};
    const cuuint32_t box[4] = {(cuuint32_t)(a_sw / 2), (cuuint32_t)p->Wp, (cuuint32_t)p->Hw, (cuuint32_t)p->TBW};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(in), gdim, gstr, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(a_sw), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (p->stride == 1) {
    const cuuint64_t gdim[4] = {C, W, H, (cuuint64_t)B};
    const cuuint64_t gstr[3] = {C * 2, W * C * 2, H * W * C * 2};
    const cuuint32_t box_v1[4] = {(cuuint32_t)p->kc, (cuuint32_t)p->TW, (cuuint32_t)p->TH, (cuuint32_t)p->TB};
    const cuuint32_t box_v4[4] = {(cuuint32_t)p->kc, (cuuint32_t)kTwWp, (cuuint32_t)kTwHp, 1};      // block + halo
    const cuuint32_t* box = p->use_tapwin ? box_v4 : box_v1;
    const cuuint32_t es[4] = {1, 1, 1, 1};
    r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(in), gdim, gstr, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(p->sw), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    // [B][H/2][2][W/2][2C]: d0 = (col parity, channel), d1 = col/2, d2 = row parity, d3 = row/2, d4 = batch
    const cuuint64_t gdim[5] = {2 * C, W / 2, 2, H / 2, (cuuint64_t)B};
    const cuuint64_t gstr[4] = {2 * C * 2, W * C * 2, 2 * W * C * 2, H * W * C * 2};
    const cuuint32_t box[5] = {(cuuint32_t)p->kc, (cuuint32_t)p->TW, 1, (cuuint32_t)p->TH, (cuuint32_t)p->TB};
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(in), gdim, gstr, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(p->sw), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(activations %dx%dx%d, B=%d, stride %d) failed with %d", p->H, p->W, p->Cin_p, B,
              p->stride, (int)r);
    return EGN_ERR_CUDA;
  }
  return EGN_OK;
}

// Function attributes (the > 48 KB dynamic shared-memory opt-in) and the SM count are PER DEVICE: a process that
// runs HC on several GPUs (torch.cuda.set_device, one thread per GPU, model.to('cuda:1')) must opt in on each.
// Returns the SM count of the current device (0 after setting the error on failure).
static int device_setup() {
  constexpr int kMaxDev = 64;
  static std::mutex mu;
  static int sms[kMaxDev] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) {
    set_error("conv_tc: cudaGetDevice failed or ordinal out of range");
    return 0;
  }
  std::lock_guard<std::mutex> lock(mu);
  if (sms[dev]) return sms[dev];
  cudaError_t e = cudaSuccess;
  auto opt_in = [&](const void* fn, int bytes) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  };
  opt_in((const void*)conv_tc_kernel<128>, 227 * 1024);
  opt_in((const void*)conv_tc_kernel<64>, 227 * 1024);
  opt_in((const void*)conv_tc_kernel<32>, 227 * 1024);
  opt_in((const void*)conv_run_kernel, 200 * 1024 + 2048);
  opt_in((const void*)conv_tapwin_kernel<128, false>, 227 * 1024);
  opt_in((const void*)conv_tapwin_kernel<64, false>, 227 * 1024);
  opt_in((const void*)conv_tapwin_kernel<128, true>, 227 * 1024);
  opt_in((const void*)conv_tapwin_kernel<64, true>, 227 * 1024);
  opt_in((const void*)conv_persist_kernel<false>, 227 * 1024);
  opt_in((const void*)conv_persist_kernel<true>, 227 * 1024);
  opt_in((const void*)conv_persist_kernel<false, true>, 227 * 1024);
  int n = 0;
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess || n <= 0) {
    set_error("conv_tc: per-device kernel setup failed on device %d: %s", dev, cudaGetErrorString(e));
    return 0;
  }
  sms[dev] = n;
  return n;
}

template <int SW>
static int launch_sw(TcConvPlan* p, const CUtensorMap& ma, const CUtensorMap& m_res, const CUtensorMap& m_out, const TcParams& tp,
                     dim3 grid, cudaStream_t st) {
  EGN_CUDA_CHECK(launch_pdl(conv_tc_kernel<SW>, grid, dim3(kTcThreads), p->smem_bytes, st, ma, p->map_b, m_res, m_out, tp));
  EGN_LAUNCH_CHECK("conv_tc_kernel");
  return EGN_OK;
}

int launch_conv_tc(TcConvPlan* p, const ConvArgs& a, cudaStream_t st) {
  if (!p) {
    set_error("launch_conv_tc: null plan");
    return EGN_ERR_STATE;
  }
  const int num_sms = device_setup();
  if (!num_sms) return EGN_ERR_CUDA;
  CUtensorMap ma;
  {
    std::lock_guard<std::mutex> lock(p->mu);
    auto key = std::make_pair(a.in, a.B);
    auto it = p->a_maps.find(key);
    if (it == p->a_maps.end()) {
      if (p->a_maps.size() > 64) p->a_maps.clear();
      CUtensorMap m;
      if (int rc = make_a_map(p, a.in, a.B, &m)) return rc;
      it = p->a_maps.emplace(key, m).first;
    }
    ma = it->second;
  }
  if (p->use_persist) {
    PersistParams pp{};
    RunParams& rp = pp.r;
    rp.B = a.B; rp.H = p->H; rp.W = p->W; rp.Cout_p = p->Cout_p; rp.Cout = p->Cout; rp.Cin_p = p->cin_a;
    rp.taps = p->ksize * p->ksize; rp.relu = a.relu;
    rp.halo = p->halo; rp.Wp = p->Wp; rp.Hw = p->Hw; rp.THW = p->THW; rp.TBW = p->TBW;
    pp.blk = p->blk; pp.BW = p->BW; pp.BH = p->BH; pp.TX = p->TX;
    pp.wins_x = p->blk ? ceil_div(p->W, p->BW) : 1;
    rp.win_per_img = p->blk ? pp.wins_x * ceil_div(p->H, p->BH) : ceil_div(p->H, p->THW);
    rp.lead = p->halo * (p->Wp + 1);
    rp.m_run = p->TBW * p->Hw * p->Wp - 2 * rp.lead;
    rp.img_rows = p->Hw * p->Wp;
    rp.inv_wp = 1.0f / (float)p->Wp;
    rp.inv_img = 1.0f / (float)rp.img_rows;
    rp.T = p->T; rp.rows_alloc = p->rows_alloc;
    rp.n_tile = p->n_tile; rp.kchunks = p->a_kchunks; rp.cin_k = p->cin_k; rp.b_stages = p->b_stages;
    rp.a_bytes = (uint32_t)(p->TBW * p->Hw * p->Wp) * (uint32_t)p->a_sw;
    rp.b_bytes = (uint32_t)p->n_tile * 128u;
    rp.tmem_cols = p->tmem_cols;
    rp.bias = a.bias;
    rp.res = static_cast<const __half*>(a.res);
    rp.out = static_cast<__half*>(a.out);
    rp.heatmap = a.heatmap; rp.xs = a.xs; rp.ys = a.ys; rp.coord_maps = a.coord_maps;
    rp.dbg = getenv("EGN_TC_DBG") ? atoi(getenv("EGN_TC_DBG")) : 0;
    pp.split = p->split ? 1 : 0;
    pp.nh = p->Cin_p / 16;
    pp.a_slots = p->a_slots;
    pp.a_sw = p->a_sw;
    pp.grouped = (getenv("EGN_TC_GROUPED") && atoi(getenv("EGN_TC_GROUPED")) == 0) ? 0 : 1;
    // L2 bulk prefetch of the residual rows by the A producer (direct, non-staged epilogue only):
    // EGN_TC_RES_PREFETCH = 2 (default) only when the output channels are not split over blockIdx.y, 1 always,
    // 0 never.  A split layer would prefetch the all-channel rows once per half -- 2x the residual DRAM traffic
    // (profiles/r01d_ncu_conv_persist_96ch.md) for no measurable gain (10.91k vs 10.92k crops/s, 5 A/B runs).
    static const int res_prefetch = getenv("EGN_TC_RES_PREFETCH") ? atoi(getenv("EGN_TC_RES_PREFETCH")) : 2;
    if (res_prefetch == 0 || (res_prefetch == 2 && p->n_tiles > 1)) rp.dbg |= 128;
    rp.ts = nullptr;
    static unsigned long long* d_ts3 = nullptr;
    if (getenv("EGN_TC_TS")) {
      if (!d_ts3) cudaMalloc(&d_ts3, 256 * 128 * sizeof(unsigned long long));
      cudaMemsetAsync(d_ts3, 0, 256 * 128 * sizeof(unsigned long long), st);
      rp.ts = d_ts3;
    }
    pp.n_windows = rp.win_per_img * ceil_div(a.B, p->TBW);
    pp.b_resident = p->b_resident;
    pp.b_rows = p->n_tile;
    pp.pack_tail = p->pack_tail ? 1 : 0;
    pp.w_tiles = p->w_tiles;
    const bool head = a.heatmap || a.coord_maps || getenv("EGN_TC_EPI_GENERIC");
    const bool pair = p->use_pair && !head && pp.n_windows >= 2;
    if (pair) rp.n_tile = 2 * p->n_tile;         // MMA / epilogue width; b_rows (and b_bytes) stay per-CTA
    // K-split accumulators (EGN_TC_KSPLIT=k, off by default): where a window is a single M tile the MMAs form one
    // dependent chain (62 SM cycles per MMA for every N <= 128, profiles/r01k_umma_rate.log, r01m_timeline_96ch.log);
    // dealing the K loop over k accumulators was measured on the 96-channel layers and did NOT pay: 85.9 -> 87.1 (k=2)
    // -> 88.3 us (k=4) with the N split, 82.1 -> 83.7 us (k=2) with CTA pairs.  Kept as a tested switch.
    {
      const int ks_env = getenv("EGN_TC_KSPLIT") ? atoi(getenv("EGN_TC_KSPLIT")) : 0;
      const int ksteps = rp.taps * ((p->Cin_p + 15) / 16);
      int ks = 1;
      if (!head && ks_env > 1) {
        ks = ks_env;
        while (ks > 1 && (2 * ks * p->T * rp.n_tile > 512 || ks > ksteps)) --ks;
      }
      if (p->split) ks = 2;          // fp16x2: accumulator H (hi*hi) and L (cross terms), summed by the epilogue
      pp.ksplit = ks;
      rp.tmem_cols = pow2_cols(2 * ks * p->T * rp.n_tile);
    }
    // chunk phasing of the single window slot (see PersistParams)
    pp.chunk_phase = (p->split && p->a_slots == 1 && p->a_kchunks >= 2 && p->a_kchunks <= 4 && p->b_resident && !pair) ? 1 : 0;
    if (p->a_slots == 1 && !pp.chunk_phase) {
      set_error("launch_conv_tc: a single window slot needs the chunk-phased ring");
      return EGN_ERR_STATE;
    }
    if (pp.chunk_phase) {
      const int groups = 2 * pp.nh, spc = p->a_sw / 32;
      int cum = 0;
      for (int c = 0; c < 4; ++c) {
        int n_c = 0;
        for (int g = 0; g < groups && c < p->a_kchunks; ++g)
          if (((g >> 1) + (g & 1) * pp.nh) / spc == c) ++n_c;
        cum += rp.taps * n_c * p->T;
        pp.phase_end[c] = cum;
      }
    }
    CUtensorMap m_res = ma, m_out = ma;          // placeholders when the epilogue is not staged
    if (p->n_stage && !head) {
      pp.n_stage = p->n_stage; pp.cb = p->cb; pp.rows_stage = p->rows_stage;
      pp.nblk_plane = pair ? 2 * p->nblk : p->nblk;
      pp.nblk = (p->split ? 2 : 1) * pp.nblk_plane;
      pp.blk_bytes = p->blk_bytes;
      std::lock_guard<std::mutex> lock(p->mu);
      for (int which = 0; which < 2; ++which) {
        const void* ptr = which ? a.out : a.res;
        if (!ptr) continue;
        auto key = std::make_pair(ptr, a.B);
        auto it = p->io_maps.find(key);
        if (it == p->io_maps.end()) {
          if (p->io_maps.size() > 64) p->io_maps.clear();
          CUtensorMap m;
          if (int rc = make_io_map(p, ptr, a.B, &m)) return rc;
          it = p->io_maps.emplace(key, m).first;
        }
        (which ? m_out : m_res) = it->second;
      }
      if (!a.res) m_res = m_out;
    }
    dim3 grid((unsigned)std::min(pp.n_windows, num_sms), (unsigned)p->n_tiles);
    if (pair) {
      // one cluster of two CTAs per SM pair; the extra bias floats of the full-width tile need 4 * n_tile more bytes
      const dim3 pgrid((unsigned)(std::min(pp.n_windows, num_sms) & ~1), 1);
      EGN_CUDA_CHECK(launch_pdl_cluster(conv_persist_kernel<false, true>, pgrid, dim3(kPersistThreads), p->pair_smem, st, 2,
                                        ma, p->map_b, m_res, m_out, pp));
    } else if (head)
      EGN_CUDA_CHECK(launch_pdl(conv_persist_kernel<true>, grid, dim3(kPersistThreadsHead), p->smem_bytes, st, ma, p->map_b, m_res, m_out, pp));
    else
      EGN_CUDA_CHECK(launch_pdl(conv_persist_kernel<false>, grid, dim3(kPersistThreads), p->smem_bytes, st, ma, p->map_b, m_res, m_out, pp));
    EGN_LAUNCH_CHECK("conv_persist_kernel");
    if (rp.ts && getenv("EGN_TC_TS_DUMP")) {
      cudaStreamSynchronize(st);
      std::vector<unsigned long long> h(256 * 128);
      cudaMemcpy(h.data(), d_ts3, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      // CTA 0 and CTA 77, first windows: times relative to the CTA's first stamp
      for (int cta : {0, 77}) {
        const unsigned long long t0 = h[(size_t)cta * 128];
        for (int j = 0; j < 6; ++j) {
          const unsigned long long* q = &h[((size_t)cta * 8 + j) * 16];
          if (!q[0]) continue;
          fprintf(stderr, "[egn-ts3] cta %d win %d: mma-loop-top %.2f acc-empty-ok %.2f a-full-ok %.2f mma-issued %.2f | epi acc-full-ok %.2f epi-done %.2f (us) | mma phase %llu SM cycles\n",
                  cta, j, (q[0] - t0) * 1e-3, (q[1] - t0) * 1e-3, (q[2] - t0) * 1e-3, (q[3] - t0) * 1e-3, (q[4] - t0) * 1e-3, (q[5] - t0) * 1e-3, q[7] - q[6]);
          if (q[12])
            fprintf(stderr, "[egn-ts3]            staged: residual landed %.2f items done (warp 6) %.2f | dma: all staged %.2f store read out %.2f (us)\n",
                    (q[12] - t0) * 1e-3, (q[13] - t0) * 1e-3, (q[14] - t0) * 1e-3, (q[15] - t0) * 1e-3);
          if (q[8])
            fprintf(stderr, "[egn-ts3]            chunk phases: A-issued %.2f a-full[1]-ok %.2f | producer: chunk0 free %.2f chunk1 free %.2f (us)\n",
                    (q[8] - t0) * 1e-3, (q[9] - t0) * 1e-3, (q[10] - t0) * 1e-3, (q[11] - t0) * 1e-3);
        }
      }
    }
    return EGN_OK;
  }
  if (p->use_run) {
    RunParams rp{};
    rp.B = a.B; rp.H = p->H; rp.W = p->W; rp.Cout_p = p->Cout_p; rp.Cout = p->Cout; rp.Cin_p = p->cin_a;
    rp.run_split = p->split ? 1 : 0; rp.run_nh = p->Cin_p / 16;
    rp.taps = p->ksize * p->ksize; rp.relu = a.relu;
    rp.halo = p->halo; rp.Wp = p->Wp; rp.Hw = p->Hw; rp.THW = p->THW; rp.TBW = p->TBW;
    rp.win_per_img = ceil_div(p->H, p->THW);
    rp.lead = p->halo * (p->Wp + 1);
    rp.m_run = p->TBW * p->Hw * p->Wp - 2 * rp.lead;
    rp.img_rows = p->Hw * p->Wp;
    rp.inv_wp = 1.0f / (float)p->Wp;
    rp.inv_img = 1.0f / (float)rp.img_rows;
    rp.T = p->T; rp.rows_alloc = p->rows_alloc;
    rp.n_tile = p->n_tile; rp.kchunks = p->kchunks; rp.cin_k = p->cin_k; rp.b_stages = p->b_stages;
    rp.a_bytes = (uint32_t)(p->TBW * p->Hw * p->Wp) * 128u;
    rp.b_bytes = (uint32_t)p->n_tile * 128u;
    rp.tmem_cols = p->tmem_cols;
    rp.bias = a.bias;
    rp.res = static_cast<const __half*>(a.res);
    rp.out = static_cast<__half*>(a.out);
    rp.heatmap = a.heatmap; rp.xs = a.xs; rp.ys = a.ys; rp.coord_maps = a.coord_maps;
    rp.dbg = getenv("EGN_TC_DBG") ? atoi(getenv("EGN_TC_DBG")) : 0;
    rp.ts = nullptr;
    static unsigned long long* d_ts = nullptr;
    const size_t n_cta = (size_t)(ceil_div(p->H, p->THW) * ceil_div(a.B, p->TBW)) * p->n_tiles;
    if (getenv("EGN_TC_TS")) {
      if (!d_ts) cudaMalloc(&d_ts, 8192 * 8 * sizeof(unsigned long long));
      if (n_cta <= 8192) rp.ts = d_ts;
    }
    dim3 grid((unsigned)(rp.win_per_img * ceil_div(a.B, p->TBW)), (unsigned)p->n_tiles);
    EGN_CUDA_CHECK(launch_pdl(conv_run_kernel, grid, dim3(kTcThreads), p->smem_bytes, st, ma, p->map_b, rp));
    EGN_LAUNCH_CHECK("conv_run_kernel");
    if (rp.ts && getenv("EGN_TC_TS_DUMP")) {
      cudaStreamSynchronize(st);
      std::vector<unsigned long long> h(n_cta * 8);
      cudaMemcpy(h.data(), d_ts, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      unsigned long long t0 = ~0ull, t1 = 0;
      double acc[7] = {0, 0, 0, 0, 0, 0, 0};
      for (size_t c = 0; c < n_cta; ++c) {
        t0 = std::min(t0, h[c * 8]);
        t1 = std::max(t1, h[c * 8 + 6]);
        for (int i = 1; i < 7; ++i) acc[i] += (double)(h[c * 8 + i] - h[c * 8]);
      }
      fprintf(stderr, "[egn-ts] ctas=%zu span=%.2fus | avg since CTA start (us): init %.2f, A-ready %.2f, mma-issued %.2f, acc-ready %.2f, epi-done %.2f, dealloc %.2f\n",
              n_cta, (t1 - t0) * 1e-3, acc[1] / n_cta * 1e-3, acc[2] / n_cta * 1e-3, acc[3] / n_cta * 1e-3, acc[4] / n_cta * 1e-3,
              acc[5] / n_cta * 1e-3, acc[6] / n_cta * 1e-3);
    }
    return EGN_OK;
  }
  if (p->use_tapwin) {
    TapWinParams wp{};
    wp.B = a.B; wp.H = p->H; wp.W = p->W; wp.Cout_p = p->Cout_p; wp.Cout = p->Cout; wp.Cin_p = p->Cin_p; wp.relu = a.relu;
    wp.tiles_x = ceil_div(p->W, 8); wp.tiles_y = ceil_div(p->H, 16);
    wp.n_tile = p->n_tile; wp.kc = p->kc; wp.kchunks = p->kchunks; wp.na = p->tw_na; wp.nb = p->tw_nb;
    wp.tap_k = p->split ? p->Cin_p : p->cin_k;
    wp.a_bytes = (uint32_t)(kTwWp * kTwHp) * (uint32_t)p->sw;
    wp.b_bytes = (uint32_t)(p->tw_pair ? p->n_tile / 2 : p->n_tile) * (uint32_t)p->sw;
    wp.tmem_cols = p->tmem_cols;
    wp.split = p->split ? 1 : 0;
    wp.nh = p->Cin_p / 16;
    wp.bias = a.bias;
    wp.res = static_cast<const __half*>(a.res);
    wp.out = static_cast<__half*>(a.out);
    wp.heatmap = a.heatmap; wp.xs = a.xs; wp.ys = a.ys; wp.coord_maps = a.coord_maps;
    wp.keep_a = (getenv("EGN_TC_KEEP_A") && atoi(getenv("EGN_TC_KEEP_A"))) ? 1 : 0;   // measured: no gain (128.6 vs 129.8 us), off
    CUtensorMap m_res = ma, m_out = ma;          // placeholders when the epilogue is not staged
    wp.staged = (p->n_stage && !a.heatmap && !a.coord_maps) ? 1 : 0;
    if (wp.staged) {
      wp.cb = p->cb; wp.nblk_plane = p->nblk; wp.blk_bytes = p->blk_bytes;
      std::lock_guard<std::mutex> lock(p->mu);
      for (int which = 0; which < 2; ++which) {
        const void* ptr = which ? a.out : a.res;
        if (!ptr) continue;
        auto key = std::make_pair(ptr, a.B);
        auto it = p->io_maps.find(key);
        if (it == p->io_maps.end()) {
          if (p->io_maps.size() > 64) p->io_maps.clear();
          CUtensorMap m;
          if (int rc = make_io_map(p, ptr, a.B, &m)) return rc;
          it = p->io_maps.emplace(key, m).first;
        }
        (which ? m_out : m_res) = it->second;
      }
      if (!a.res) m_res = m_out;
    }
    wp.total_tiles = wp.tiles_x * wp.tiles_y * a.B;
    wp.grouped = (getenv("EGN_TC_GROUPED") && atoi(getenv("EGN_TC_GROUPED")) == 0) ? 0 : 1;
    const bool persist4 = p->tw_persist && wp.staged;
    wp.acc_sets = persist4 ? 2 : 1;
    wp.stage_off = persist4 ? p->tw_stage_off : 0u;
    wp.stage_bytes = p->tw_stage_bytes;            // (the barrier block sits behind the region either way)
    if (p->tw_persist && !persist4) wp.tmem_cols = p->tmem_cols;      // head extras: one tile per CTA, direct epilogue
    wp.n_fold = p->tw_fold;
    const int units1 = wp.total_tiles * wp.n_fold;           // work units: (tile, N part)
    dim3 grid((unsigned)(persist4 ? std::min(units1, num_sms) : units1), (unsigned)p->n_tiles);
    if (p->tw_pair) {
      // clusters of two CTAs along x: tiles 2q, 2q + 1 per pair (an odd last tile gets a past-the-end partner)
      const int units = (wp.total_tiles + 1) / 2 * wp.n_fold;
      grid.x = 2u * (unsigned)(persist4 ? std::min(units, num_sms / 2) : units);
    }
    static unsigned long long* d_ts4 = nullptr;
    const size_t n_cta = (size_t)grid.x * grid.y;
    if (getenv("EGN_TC_TS") && n_cta <= 8192) {
      if (!d_ts4) cudaMalloc(&d_ts4, 8192 * 8 * sizeof(unsigned long long));
      cudaMemsetAsync(d_ts4, 0, 8192 * 8 * sizeof(unsigned long long), st);
      wp.ts = d_ts4;
    }
    if (p->tw_pair) {
      if (p->sw == 128)
        EGN_CUDA_CHECK(launch_pdl_cluster(conv_tapwin_kernel<128, true>, grid, dim3(kTcThreads), p->smem_bytes, st, 2, ma, p->map_b, m_res, m_out, wp));
      else
        EGN_CUDA_CHECK(launch_pdl_cluster(conv_tapwin_kernel<64, true>, grid, dim3(kTcThreads), p->smem_bytes, st, 2, ma, p->map_b, m_res, m_out, wp));
    } else if (p->sw == 128)
      EGN_CUDA_CHECK(launch_pdl(conv_tapwin_kernel<128, false>, grid, dim3(kTcThreads), p->smem_bytes, st, ma, p->map_b, m_res, m_out, wp));
    else
      EGN_CUDA_CHECK(launch_pdl(conv_tapwin_kernel<64, false>, grid, dim3(kTcThreads), p->smem_bytes, st, ma, p->map_b, m_res, m_out, wp));
    EGN_LAUNCH_CHECK("conv_tapwin_kernel");
    if (wp.ts && getenv("EGN_TC_TS_DUMP")) {
      cudaStreamSynchronize(st);
      std::vector<unsigned long long> h(n_cta * 8);
      cudaMemcpy(h.data(), d_ts4, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      unsigned long long t0 = ~0ull, t1 = 0;
      double acc[7] = {0, 0, 0, 0, 0, 0, 0};
      for (size_t c = 0; c < n_cta; ++c) {
        t0 = std::min(t0, h[c * 8]);
        t1 = std::max(t1, h[c * 8 + 6]);
        for (int i = 1; i < 7; ++i) acc[i] += (double)(h[c * 8 + i] - h[c * 8]);
      }
      fprintf(stderr, "[egn-ts4] ctas=%zu span=%.2fus | avg since CTA start (us): init %.2f, first operands %.2f, mma-issued %.2f, acc-ready %.2f, epi-done %.2f, dealloc %.2f\n",
              n_cta, (t1 - t0) * 1e-3, acc[1] / n_cta * 1e-3, acc[2] / n_cta * 1e-3, acc[3] / n_cta * 1e-3, acc[4] / n_cta * 1e-3,
              acc[5] / n_cta * 1e-3, acc[6] / n_cta * 1e-3);
      // the first CTAs in launch order, relative to the kernel's first stamp
      for (size_t c = 0; c < 4 && c < n_cta; ++c)
        fprintf(stderr, "[egn-ts4] cta %zu: start %.2f init %.2f first-operands %.2f mma-issued %.2f acc-ready %.2f epi-done %.2f dealloc %.2f (us)\n", c,
                (h[c * 8] - t0) * 1e-3, (h[c * 8 + 1] - t0) * 1e-3, (h[c * 8 + 2] - t0) * 1e-3, (h[c * 8 + 3] - t0) * 1e-3,
                (h[c * 8 + 4] - t0) * 1e-3, (h[c * 8 + 5] - t0) * 1e-3, (h[c * 8 + 6] - t0) * 1e-3);
    }
    return EGN_OK;
  }
  TcParams tp{};
  tp.B = a.B; tp.OH = p->OH; tp.OW = p->OW; tp.Cout_p = p->Cout_p; tp.Cout = p->Cout; tp.Cin_p = p->Cin_p;
  tp.split = p->split ? 1 : 0; tp.cin_a = p->cin_a; tp.nh = p->Cin_p / 16;
  tp.keep_a = (getenv("EGN_TC_KEEP_A") && atoi(getenv("EGN_TC_KEEP_A"))) ? 1 : 0;
  tp.grouped = (getenv("EGN_TC_GROUPED") && atoi(getenv("EGN_TC_GROUPED")) == 0) ? 0 : 1;
  tp.taps = p->ksize * p->ksize; tp.ksize = p->ksize; tp.stride = p->stride; tp.pad = p->pad; tp.relu = a.relu;
  tp.TW = p->TW; tp.TH = p->TH; tp.TB = p->TB;
  tp.tiles_w = ceil_div(p->OW, p->TW);
  tp.tiles_h = ceil_div(p->OH, p->TH);
  tp.n_tile = p->n_tile; tp.kc = p->kc; tp.kchunks = p->kchunks; tp.cin_k = p->cin_k; tp.stages = p->stages;
  tp.a_bytes = (uint32_t)(p->TW * p->TH * p->TB) * (uint32_t)p->sw;
  tp.b_bytes = (uint32_t)p->n_tile * (uint32_t)p->sw;
  tp.tmem_cols = p->tmem_cols;
  tp.bias = a.bias;
  tp.res = static_cast<const __half*>(a.res);
  tp.out = static_cast<__half*>(a.out);
  tp.heatmap = a.heatmap; tp.xs = a.xs; tp.ys = a.ys; tp.coord_maps = a.coord_maps;
  CUtensorMap m_res = ma, m_out = ma;          // placeholders when the epilogue is not staged
  tp.staged = (p->n_stage && !a.heatmap && !a.coord_maps) ? 1 : 0;
  if (tp.staged) {
    tp.cb = p->cb; tp.nblk_plane = p->nblk; tp.blk_bytes = p->blk_bytes;
    std::lock_guard<std::mutex> lock(p->mu);
    for (int which = 0; which < 2; ++which) {
      const void* ptr = which ? a.out : a.res;
      if (!ptr) continue;
      auto key = std::make_pair(ptr, a.B);
      auto it = p->io_maps.find(key);
      if (it == p->io_maps.end()) {
        if (p->io_maps.size() > 64) p->io_maps.clear();
        CUtensorMap m;
        if (int rc = make_io_map(p, ptr, a.B, &m)) return rc;
        it = p->io_maps.emplace(key, m).first;
      }
      (which ? m_out : m_res) = it->second;
    }
    if (!a.res) m_res = m_out;
  }
  dim3 grid((unsigned)(tp.tiles_w * tp.tiles_h * ceil_div(a.B, p->TB)), (unsigned)p->n_tiles);
  switch (p->sw) {
    case 128: return launch_sw<128>(p, ma, m_res, m_out, tp, grid, st);
    case 64: return launch_sw<64>(p, ma, m_res, m_out, tp, grid, st);
    default: return launch_sw<32>(p, ma, m_res, m_out, tp, grid, st);
  }
}

}  // namespace egn

// Plan of one fused conv as text (debug / tests): which kernel runs the shape and how it is tiled.  Geometry only:
// works without a GPU (tc_conv_plan_create with no weights).
extern "C" int egn_debug_conv_plan(int dtype, int Cin, int Cout, int H, int W, int ksize, int stride, char* out, int out_len) {
  using namespace egn;
  EGN_REQUIRE(out && out_len > 0, "egn_debug_conv_plan: null output");
  EGN_REQUIRE((dtype == 1 || dtype == 2) && Cin > 0 && Cout > 0 && H > 0 && W > 0 && (ksize == 1 || ksize == 3) &&
                  (stride == 1 || stride == 2),
              "egn_debug_conv_plan: bad arguments");
  ConvArgs a{};
  a.B = 1; a.H = H; a.W = W;
  a.Cin_p = round_up(Cin, 16);           // channel padding of the engine (hrnet_graph.h kChanAlign)
  a.Cout_p = round_up(Cout, 16);
  a.Cout = Cout;
  a.ksize = ksize; a.stride = stride; a.pad = ksize == 3 ? 1 : 0;
  a.OH = (H + 2 * a.pad - ksize) / stride + 1;
  a.OW = (W + 2 * a.pad - ksize) / stride + 1;
  a.split = dtype == 2 ? 1 : 0;
  if (!tc_conv_supported(a)) {
    snprintf(out, (size_t)out_len, "kernel=unsupported");
    return EGN_OK;
  }
  TcConvPlan* p = nullptr;
  if (int rc = tc_conv_plan_create(a, nullptr, &p)) return rc;
  const char* kernel = p->use_persist ? "v3-persist" : (p->use_run ? "v2-run" : (p->use_tapwin ? "v4-tapwin" : "v1-tap"));
  if (p->use_persist)
    snprintf(out, (size_t)out_len, "kernel=%s n_tile=%d n_tiles=%d blk=%dx%d a_sw=%d a_slots=%d T=%d resident=%d stage=%d pair=%d smem=%zu tmem=%u",
             kernel, p->n_tile, p->n_tiles, p->blk ? p->BW : 0, p->blk ? p->BH : 0, p->a_sw, p->a_slots, p->T, p->b_resident, p->n_stage,
             p->use_pair ? 1 : 0, p->smem_bytes, p->tmem_cols);
  else if (p->use_tapwin)
    snprintf(out, (size_t)out_len, "kernel=%s n_tile=%d n_tiles=%d sw=%d kchunks=%d na=%d nb=%d staged=%d persist=%d pair=%d fold=%d smem=%zu tmem=%u",
             kernel, p->n_tile, p->n_tiles, p->sw, p->kchunks, p->tw_na, p->tw_nb, p->n_stage, p->tw_persist ? 1 : 0, p->tw_pair ? 1 : 0,
             p->tw_fold, p->smem_bytes, p->tmem_cols);
  else
    snprintf(out, (size_t)out_len, "kernel=%s n_tile=%d n_tiles=%d sw=%d kchunks=%d stages=%d staged=%d T=%d smem=%zu tmem=%u", kernel,
             p->n_tile, p->n_tiles, p->sw, p->kchunks, p->stages, p->n_stage, p->T, p->smem_bytes, p->tmem_cols);
  tc_conv_plan_destroy(p);
  return EGN_OK;
}

// ---------------------------------------------------------------------------
// Hardware probe: does a K-major swizzled UMMA descriptor whose start address is offset by an
// arbitrary number of ROWS (not a multiple of the 8-row swizzle atom) address the rows TMA wrote?
// D[i][n] = sum_k A[off+i][k] * B[n][k] with B = identity, for a few encodings of `base_offset`.
// Used once to validate the flattened-run conv kernel's operand addressing.
// ---------------------------------------------------------------------------
namespace egn {

template <int SW>
__global__ void __launch_bounds__(128)
umma_probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int row_off,
                  int bo_mode, float* __restrict__ out /* [128][SW/2] */) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int KC = SW / 2;
  uint8_t* sa = smem;                       // 256 rows x SW bytes
  uint8_t* sb = smem + 256 * SW;            // KC rows x SW bytes (1024-aligned since 256*SW is)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + 8192);
  uint64_t* bar2 = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar2, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(slot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 256 * SW + KC * SW);
    tma_load_2d(sa, &map_a, bar, 0, 0);
    tma_load_2d(sb, &map_b, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t start = smem_u32(sa) + (uint32_t)row_off * SW;
    uint64_t adesc = make_smem_desc(start, SW);
    // bo_mode >= 16: the 8-row groups of the operand are (bo_mode >> 4) rows apart instead of 8 (stride byte offset
    // = rows * SW): row i of the tile is operand row  off + (i / 8) * rows + i % 8  (block-shaped conv windows)
    if (bo_mode >= 16) {
      const uint64_t sbo = (uint64_t)(((uint32_t)(bo_mode >> 4) * SW) >> 4);
      adesc = (adesc & ~((uint64_t)0x3FFF << 32)) | (sbo << 32);
      bo_mode &= 15;
    }
    uint32_t bo = 0;
    if (bo_mode == 1) bo = (uint32_t)row_off & 7u;          // row phase inside the 8-row atom
    if (bo_mode == 2) bo = (start >> 7) & 7u;               // literal (addr >> 7) & 7
    adesc |= (uint64_t)bo << 49;
    const uint64_t bdesc = make_smem_desc(smem_u32(sb), SW);
    const uint32_t idesc = (1u << 4) | ((uint32_t)(KC >> 3) << 17) | ((128u >> 4) << 24);
#pragma unroll
    for (int k = 0; k < SW / 32; ++k) umma_f16(tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
    umma_commit(bar2);
  }
  __syncthreads();
  mbar_wait(bar2, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < KC; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[row * KC + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

template <int SW>
static int run_probe(int row_off, int bo_mode, const __half* d_a, const __half* d_b, float* d_out, cudaStream_t st) {
  constexpr int KC = SW / 2;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled unavailable");
    return EGN_ERR_CUDA;
  }
  CUtensorMap ma, mb;
  const cuuint32_t es[2] = {1, 1};
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)KC, 256};
    const cuuint64_t gstr[1] = {(cuuint64_t)KC * 2};
    const cuuint32_t box[2] = {(cuuint32_t)KC, 256};
    if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(d_a), gdim, gstr, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(SW), CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_error("probe: encode A failed");
      return EGN_ERR_CUDA;
    }
  }
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)KC, (cuuint64_t)KC};
    const cuuint64_t gstr[1] = {(cuuint64_t)KC * 2};
    const cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)KC};
    if (enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(d_b), gdim, gstr, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(SW), CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_error("probe: encode B failed");
      return EGN_ERR_CUDA;
    }
  }
  const size_t smem = 1024 + 256 * SW + 8192 + 64;
  EGN_CUDA_CHECK(cudaFuncSetAttribute(umma_probe_kernel<SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_probe_kernel<SW><<<1, 128, smem, st>>>(ma, mb, row_off, bo_mode, d_out);
  EGN_LAUNCH_CHECK("umma_probe_kernel");
  EGN_CUDA_CHECK(cudaStreamSynchronize(st));
  return EGN_OK;
}

}  // namespace egn

extern "C" int egn_debug_umma_probe(int swizzle_bytes, int row_off, int bo_mode, const void* a_f16, const void* b_f16,
                                    float* out, void* stream) {
  using namespace egn;
  EGN_REQUIRE(a_f16 && b_f16 && out, "egn_debug_umma_probe: null pointer");
  EGN_REQUIRE(row_off >= 0 && row_off <= 128, "egn_debug_umma_probe: row_off out of range");
  EGN_REQUIRE(bo_mode < 16 || row_off + 15 * (bo_mode >> 4) + 8 <= 256, "egn_debug_umma_probe: group stride out of range");
  if (int rc = require_device()) return rc;
  const __half* a = static_cast<const __half*>(a_f16);
  const __half* b = static_cast<const __half*>(b_f16);
  switch (swizzle_bytes) {
    case 128: return run_probe<128>(row_off, bo_mode, a, b, out, as_stream(stream));
    case 64: return run_probe<64>(row_off, bo_mode, a, b, out, as_stream(stream));
    case 32: return run_probe<32>(row_off, bo_mode, a, b, out, as_stream(stream));
  }
  set_error("egn_debug_umma_probe: swizzle must be 32, 64 or 128");
  return EGN_ERR_INVALID;
}

// ---------------------------------------------------------------------------
// Hardware probe: issue rate of tcgen05.mma (M=128, K=16, kind::f16, SS) as a function of N and of the
// number of accumulators rotated over.  Operands are whatever is in shared memory (zeros).
// ---------------------------------------------------------------------------
namespace egn {
__global__ void __launch_bounds__(128)
umma_rate_kernel(int n, int nacc, int iters, int a_rows_shift, long long* __restrict__ out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                 // 1024 rows x 128 B
  uint8_t* sb = smem + 1024 * 128;    // 256 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + 256 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  // a_rows_shift >= 1000: fill the operands with pseudo-random finite fp16 values instead of zeros;
  // a_rows_shift >= 2000: additionally TWO issuing threads (warps 1 and 2), each driving nacc / 2 accumulators
  const bool two = a_rows_shift >= 2000;
  if (two) a_rows_shift -= 1000;
  const bool random_fill = a_rows_shift >= 1000;
  if (random_fill) a_rows_shift -= 1000;
  for (int i = threadIdx.x; i < (1024 + 256) * 128 / 4; i += 128) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15;
    reinterpret_cast<uint32_t*>(smem)[i] = random_fill ? ((h & 0x83FF83FFu) | 0x38003800u) : 0u;
  }
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *slot, 0);
  if (warp == 1 || (two && warp == 2)) {
    const int who = warp - 1;
    const int my_acc = two ? nacc / 2 : nacc;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t desc_hi = make_smem_desc(0, 128);
    const uint64_t ad0 = desc_hi | (uint64_t)(((smem_u32(sa) + (uint32_t)a_rows_shift * 128u + (uint32_t)who * my_acc * 16384u) & 0x3FFFFu) >> 4);
    const uint64_t bd0 = desc_hi | (uint64_t)((smem_u32(sb) & 0x3FFFFu) >> 4);
    long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
        for (int k = 0; k < 4; ++k) {
          uint64_t ad = ad0 + (uint64_t)(2 * k);
          uint32_t d = tmem + (uint32_t)(who * my_acc * n);
          for (int t = 0; t < my_acc; ++t, ad += 1024, d += (uint32_t)n) umma_f16(d, ad, bd0 + (uint64_t)(2 * k), idesc, 1u);
        }
      }
      umma_commit(&bar[who]);
    }
    __syncwarp();
    mbar_wait(&bar[who], 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) {
      if (!two) out_cycles[blockIdx.x] = t1 - t0;
      else atomicMax((unsigned long long*)&out_cycles[blockIdx.x], (unsigned long long)(t1 - t0));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
}  // namespace egn

extern "C" int egn_debug_umma_rate(int n, int nacc, int iters, int a_rows_shift, int ctas, double* cycles_per_mma) {
  using namespace egn;
  EGN_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && nacc >= 1 && nacc * n <= 512 && iters > 0 && ctas >= 1 && ctas <= 1024,
              "egn_debug_umma_rate: bad arguments");
  if (int rc = require_device()) return rc;
  long long* d = nullptr;
  EGN_CUDA_CHECK(cudaMalloc(&d, ctas * sizeof(long long)));
  EGN_CUDA_CHECK(cudaMemset(d, 0, ctas * sizeof(long long)));
  const size_t smem = 1024 + (1024 + 256) * 128 + 64;
  EGN_CUDA_CHECK(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_rate_kernel<<<ctas, 128, smem>>>(n, nacc, iters, a_rows_shift, d);
  EGN_LAUNCH_CHECK("umma_rate_kernel");
  EGN_CUDA_CHECK(cudaDeviceSynchronize());
  std::vector<long long> h(ctas);
  cudaMemcpy(h.data(), d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  double acc = 0;
  for (long long v : h) acc += (double)v;
  *cycles_per_mma = acc / ctas / ((double)iters * 4 * nacc);
  return EGN_OK;
}

// ---------------------------------------------------------------------------
// Hardware probe: cost of the MMA SEQUENCES the fp16x2 kernels issue (no TMA, no epilogue, operands resident):
//   pattern 0: N = n into one accumulator;  1: stacked pair per K16 slice, full-width N = 2n -> [H | L] then
//   half-width N = n -> L;  2: three-MMA form (H, L, L), N = n each;  3: as 1 but grouped per two slices (two
//   full-width, then two half-width).  The A descriptor uses `a_sw`-byte rows with its 8-row groups `sbo_rows`
//   rows apart (8 = the plain atom, 10 / 18 = block-shaped windows) and a start row that walks over the nine tap
//   shifts, as in the conv kernels.  Returns SM cycles per K16 slice.
// ---------------------------------------------------------------------------
namespace egn {
__global__ void __launch_bounds__(128)
umma_seq_kernel(int n, int pattern, int iters, int a_sw, int sbo_rows, long long* __restrict__ out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                 // 640 rows x 128 B
  uint8_t* sb = smem + 640 * 128;     // 512 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + 512 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < (640 + 512) * 128 / 4; i += 128) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15;
    reinterpret_cast<uint32_t*>(smem)[i] = (h & 0x83FF83FFu) | 0x38003800u;
  }
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *slot, 0);
  if (warp == 1) {
    const uint32_t idesc_n = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc_w = (1u << 4) | ((uint32_t)(n >> 2) << 17) | ((128u >> 4) << 24);
    const uint64_t desc_b = make_smem_desc(0, 128) | (uint64_t)((smem_u32(sb) & 0x3FFFFu) >> 4);
    const uint64_t desc_a_hi = (make_smem_desc(0, (uint32_t)a_sw) & ~((uint64_t)0x3FFF << 32)) |
                               ((uint64_t)(((uint32_t)sbo_rows * (uint32_t)a_sw) >> 4) << 32);
    const uint32_t a0 = smem_u32(sa);
    const int spc = a_sw / 32;                         // K16 slices per row
    long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
        const int tap = it % 9, r = tap / 3, q = tap - 3 * r;
        const uint32_t row = (uint32_t)(r * sbo_rows + q);
        const uint64_t ad = desc_a_hi | (uint64_t)(((a0 + row * (uint32_t)a_sw) & 0x3FFFFu) >> 4);
        const uint64_t ad_lo = ad + (uint64_t)((320u * 128u) >> 4);       // the "lo plane" box
        const uint64_t bd_lo = desc_b + (uint64_t)((256u * 128u) >> 4);
        for (int k = 0; k < 2; ++k) {
          const uint64_t ko = (uint64_t)(2 * (k % spc));
          if (pattern == 0) {
            umma_f16(tmem, ad + ko, desc_b + ko, idesc_n, 1u);
          } else if (pattern == 1) {
            umma_f16(tmem, ad + ko, desc_b + ko, idesc_w, 1u);
            umma_f16(tmem + (uint32_t)n, ad_lo + ko, desc_b + ko, idesc_n, 1u);
          } else if (pattern == 2) {
            umma_f16(tmem, ad + ko, desc_b + ko, idesc_n, 1u);
            umma_f16(tmem + (uint32_t)n, ad + ko, bd_lo + ko, idesc_n, 1u);
            umma_f16(tmem + (uint32_t)n, ad_lo + ko, desc_b + ko, idesc_n, 1u);
          }
        }
        if (pattern == 3) {
          for (int k = 0; k < 2; ++k) umma_f16(tmem, ad + (uint64_t)(2 * (k % spc)), desc_b + (uint64_t)(2 * (k % spc)), idesc_w, 1u);
          for (int k = 0; k < 2; ++k) umma_f16(tmem + (uint32_t)n, ad_lo + (uint64_t)(2 * (k % spc)), desc_b + (uint64_t)(2 * (k % spc)), idesc_n, 1u);
        }
      }
      umma_commit(&bar[0]);
    }
    __syncwarp();
    mbar_wait(&bar[0], 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out_cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
}  // namespace egn

extern "C" int egn_debug_umma_seq(int n, int pattern, int iters, int a_sw, int sbo_rows, int ctas, double* cycles_per_slice) {
  using namespace egn;
  EGN_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && pattern >= 0 && pattern <= 3 && iters > 0 && ctas >= 1 && ctas <= 1024 &&
                  (a_sw == 128 || a_sw == 64) && sbo_rows >= 8 && sbo_rows <= 18 && (pattern == 0 || pattern == 2 || 2 * n <= 256) &&
                  2 * n <= 512 && cycles_per_slice,
              "egn_debug_umma_seq: bad arguments");
  if (int rc = require_device()) return rc;
  long long* d = nullptr;
  EGN_CUDA_CHECK(cudaMalloc(&d, ctas * sizeof(long long)));
  EGN_CUDA_CHECK(cudaMemset(d, 0, ctas * sizeof(long long)));
  const size_t smem = 1024 + (640 + 512) * 128 + 64;
  EGN_CUDA_CHECK(cudaFuncSetAttribute(umma_seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_seq_kernel<<<ctas, 128, smem>>>(n, pattern, iters, a_sw, sbo_rows, d);
  EGN_LAUNCH_CHECK("umma_seq_kernel");
  EGN_CUDA_CHECK(cudaDeviceSynchronize());
  std::vector<long long> h(ctas);
  cudaMemcpy(h.data(), d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  double acc = 0;
  for (long long v : h) acc += (double)v;
  *cycles_per_slice = acc / ctas / ((double)iters * 2);
  return EGN_OK;
}
