// tsp_device.cuh — device-side arithmetic shared by every kernel of libtspb200.
//
// Exact ("reference") distances are evaluated in FP64 with explicitly rounded intrinsics
// (__dmul_rn/__dadd_rn/__dsqrt_rn/__ddiv_rn) so that nvcc can never contract a mul+add into an FMA:
// the reference is an x86-64 SSE2 build where every double operation is rounded separately
// (reference src/distutil.c:4-92).  FP32 is only ever used as a FILTER (2-opt kernels) or together
// with a guard band that falls back to FP64 (matrix kernel); see DESIGN.md §3.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tspb {

// Numeric values of the reference's weight_type enum (reference include/utility.h:45-52).
enum Metric : int { M_EUC_2D = 0, M_MAX_2D = 1, M_MAN_2D = 2, M_CEIL_2D = 3, M_GEO = 4, M_ATT = 5 };

// Padding value for "no edge": far outside any real distance, still finite in FP32 sums.
#define TSPB_BIG 1.0e30f

__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));  // MUFU.SQRT
    return r;
}

// ---- packed FP32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2, one issue slot for two values) --------------
// Packed values live in 64-bit registers for their whole life so that ptxas keeps them in aligned pairs.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2pack(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float f2lo(f32x2 v) {
    float lo;
    asm("{ .reg .b32 t; mov.b64 {%0,t}, %1; }" : "=f"(lo) : "l"(v));
    return lo;
}
__device__ __forceinline__ float f2hi(f32x2 v) {
    float hi;
    asm("{ .reg .b32 t; mov.b64 {t,%0}, %1; }" : "=f"(hi) : "l"(v));
    return hi;
}
__device__ __forceinline__ f32x2 f2sub(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 f2add(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 f2mul(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 f2fma(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// reference src/distutil.c:4-6 nint(): (long)(x + 0.5)
__device__ __forceinline__ long long nint_ref(double v) { return (long long)__dadd_rn(v, 0.5); }

// reference src/distutil.c:13-18 calc_euc2d (integer == 1)
__device__ __forceinline__ long long exact_euc(double ax, double ay, double bx, double by) {
    double ex = __dsub_rn(ax, bx), ey = __dsub_rn(ay, by);
    double len = __dsqrt_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)));
    return nint_ref(len);
}

// reference src/distutil.c:47-49 calc_ceil2d
__device__ __forceinline__ long long exact_ceil(double ax, double ay, double bx, double by) {
    double ex = __dsub_rn(ax, bx), ey = __dsub_rn(ay, by);
    double len = __dsqrt_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)));
    return (long long)ceil(len);
}

// reference src/distutil.c:20-31 calc_pseudo_euc (integer == 1)
__device__ __forceinline__ long long exact_att(double ax, double ay, double bx, double by) {
    double ex = __dsub_rn(ax, bx), ey = __dsub_rn(ay, by);
    double r = __dsqrt_rn(__ddiv_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), 10.0));
    double t = (double)nint_ref(r);
    return (long long)((t < r) ? __dadd_rn(t, 1.0) : t);
}

// reference src/distutil.c:33-37 calc_man2d: dy = fabs(p2.y - p2.y) == 0 (the bug is the spec)
__device__ __forceinline__ long long exact_man(double ax, double bx, double by) {
    double ex = fabs(__dsub_rn(ax, bx));
    double ey = fabs(__dsub_rn(by, by));
    return nint_ref(__dadd_rn(ex, ey));
}

// reference src/distutil.c:39-45 calc_max2d (same bug); dmax = utility.c:12-14
__device__ __forceinline__ long long exact_max(double ax, double bx, double by) {
    double ex = (double)nint_ref(fabs(__dsub_rn(ax, bx)));
    double ey = (double)nint_ref(fabs(__dsub_rn(by, by)));
    return (long long)(ex > ey ? ex : ey);
}

// reference src/distutil.c:51-58 calc_lat_lon for one coordinate; PI from include/distutil.h:6
__device__ __forceinline__ double geo_radians(double v) {
    const double PI = 3.14159265358979323846264;
    double deg = (double)(long long)v;
    double mn = __dsub_rn(v, deg);
    // PI * (deg + 5.0 * min / 3.0) / 180.0, evaluated left to right like C
    double inner = __dadd_rn(deg, __ddiv_rn(__dmul_rn(5.0, mn), 3.0));
    return __ddiv_rn(__dmul_rn(PI, inner), 180.0);
}

// reference src/distutil.c:60-71 calc_geo (integer == 1); inputs are (lat,lon) in radians.
// the value handed to nint(): RRR * acos(...) + 1.0
__device__ __forceinline__ double exact_geo_len(double lat1, double lon1, double lat2, double lon2) {
    const double EARTH_RAD = 6378.388;  // include/distutil.h:7
    double q1 = cos(__dsub_rn(lon1, lon2));
    double q2 = cos(__dsub_rn(lat1, lat2));
    double q3 = cos(__dadd_rn(lat1, lat2));
    double a = __dmul_rn(__dadd_rn(1.0, q1), q2);
    double b = __dmul_rn(__dsub_rn(1.0, q1), q3);
    return __dadd_rn(__dmul_rn(EARTH_RAD, acos(__dmul_rn(0.5, __dsub_rn(a, b)))), 1.0);
}
__device__ __forceinline__ long long exact_geo(double lat1, double lon1, double lat2, double lon2) {
    return nint_ref(exact_geo_len(lat1, lon1, lat2, lon2));
}
// GEO is the one metric whose value depends on library functions (cos, acos): CUDA's and glibc's are both within an ulp or
// two of the true value but not bit-identical, so an entry whose len + 0.5 lies extremely close to an integer could round
// differently on the host.  The matrix kernel counts such entries (tspb200_get_info "geo_near_boundary"; 0 on every GEO
// instance of the reference, tests/test_gpu_parity.py); when it is not 0 the caller should recompute those few distances
// with the host libm before trusting the last digit.  Window: an error of a few 1e-16 in the acos argument becomes
// ~1e-12 km for ordinary distances (window 1e-9, three decades of margin) but grows like 1 / angle for neighbouring cities
// (acos is ill-conditioned near 1): below 10 km the window is 1e-6.
__device__ __forceinline__ bool geo_near_boundary(double len) {
    const double eps = len < 10.0 ? 1.0e-6 : 1.0e-9;
    const double t = len + 0.5;
    const double fr = t - floor(t);
    return fr < eps || fr > 1.0 - eps;
}

// reference src/distutil.c:73-92 calc_dist dispatch. `pt` holds raw TSPLIB coordinates for every metric
// except GEO, where it holds the (lat,lon) radians precomputed per node by geo_radians().
// Unknown weight types fall through to EUC_2D like the reference.
__device__ __forceinline__ long long exact_dist(int metric, double2 a, double2 b) {
    switch (metric) {
        case M_ATT: return exact_att(a.x, a.y, b.x, b.y);
        case M_MAN_2D: return exact_man(a.x, b.x, b.y);
        case M_MAX_2D: return exact_max(a.x, b.x, b.y);
        case M_CEIL_2D: return exact_ceil(a.x, a.y, b.x, b.y);
        case M_GEO: return exact_geo(a.x, a.y, b.x, b.y);
        default: return exact_euc(a.x, a.y, b.x, b.y);
    }
}

// ---- move keys ---------------------------------------------------------------------------------
// A best-improvement candidate: exact integer delta and the node pair i<j.  Ordering = (delta, i, j)
// ascending, i.e. the reference's strict '<' scan in row-major (i,j) order keeps the lowest (i,j)
// among equal deltas (reference src/tabusearch.c:126-156).  "No move" is {0, INT_MAX, INT_MAX}:
// only strictly negative deltas can beat it, like mindelta = 0.
struct __align__(16) MoveKey {
    int delta;
    int i;
    int j;
    int pad;
};

__device__ __forceinline__ MoveKey key_none() {
    MoveKey k;
    k.delta = 0; k.i = 0x7fffffff; k.j = 0x7fffffff; k.pad = 0;
    return k;
}

// L2-coherent load of a key written by another block (bypasses L1).
__device__ __forceinline__ MoveKey key_load_cg(const MoveKey *p) {
    int4 v = __ldcg(reinterpret_cast<const int4 *>(p));
    MoveKey k;
    k.delta = v.x; k.i = v.y; k.j = v.z; k.pad = v.w;
    return k;
}

__device__ __forceinline__ bool key_less(const MoveKey &a, const MoveKey &b) {
    if (a.delta != b.delta) return a.delta < b.delta;
    if (a.i != b.i) return a.i < b.i;
    return a.j < b.j;
}

__device__ __forceinline__ MoveKey key_shfl_xor(const MoveKey &k, int lane_mask) {
    MoveKey o;
    o.delta = __shfl_xor_sync(0xffffffffu, k.delta, lane_mask);
    o.i = __shfl_xor_sync(0xffffffffu, k.i, lane_mask);
    o.j = __shfl_xor_sync(0xffffffffu, k.j, lane_mask);
    o.pad = 0;
    return o;
}

__device__ __forceinline__ MoveKey key_warp_min(MoveKey k) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        MoveKey o = key_shfl_xor(k, m);
        if (key_less(o, k)) k = o;
    }
    return k;
}

// 62-bit packing of a key, ordered like key_less(): valid for n <= 2^17 and |delta| < 2^27.  Used for the per-pass
// 64-bit atomicMin of the scan kernel and for the multi-GPU exchange (the two free top bits leave room for the
// generation bits of the peer-memory exchange word, see xchg_min()).
constexpr int KEY_PACK_MAX_N = 1 << 17;
constexpr int KEY_PACK_MAX_DELTA = 1 << 27;
__host__ __device__ __forceinline__ unsigned long long key_pack(int delta, int i, int j) {
    return ((unsigned long long)(unsigned)(delta + KEY_PACK_MAX_DELTA) << 34) | ((unsigned long long)(unsigned)i << 17) |
           (unsigned long long)(unsigned)j;
}
__host__ __device__ __forceinline__ void key_unpack(unsigned long long p, int *delta, int *i, int *j) {
    *delta = (int)(p >> 34) - KEY_PACK_MAX_DELTA;
    *i = (int)((p >> 17) & 0x1ffffu);
    *j = (int)(p & 0x1ffffu);
}
// "no improving move": larger than every key with delta < 0
#define KEY_PACK_NONE (key_pack(0, 0x1ffff, 0x1ffff))

// ---- programmatic dependent launch (sm_90+): the next kernel of the stream may be scheduled while this one still runs;
// everything it does before pdl_wait() must not depend on this kernel's results.  Without the launch attribute both
// are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Launch helper: `kern<<<grid, block, smem, st>>>(args...)`, optionally with the programmatic-stream-serialization
// attribute (the kernel must then call pdl_wait() before touching anything the previous kernel of the stream wrote).
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_maybe_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                           Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

// ---- TMA (bulk async copy) + mbarrier, raw PTX --------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine (SASS: UBLKCP); bytes % 16 == 0, 16-B aligned.
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace tspb
