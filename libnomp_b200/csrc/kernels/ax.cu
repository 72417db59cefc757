// Spectral-element local Poisson operator  w = A_L u  on (N+1)^3 hexahedral elements, fp64 (Nekbone ax_e /
// CEED BK5 form; definition in include/nompk.h and SURVEY.md 8(a-17)).  Not present in the reference
// (reference tests/sem.py:10-36 only tags loops; the structural precursor is the element/dof kernel of
// reference tests/nomp-api-400-impl.h:254-291 with one block per element and private temporaries).
//
// Roofline: 64 algorithmic B/DOF (u 8 + six geometric factors 48 + w 8; D is 8 n^2 bytes, constant) against
// 12n+15 flop/DOF (111 at n = 8)  ->  1.7 flop/B, HBM-bound on B200 (fp64 ridge ~5 flop/B).  At the measured
// copy bandwidth one SM has ~2.8 clocks per DOF, which allows ~350 B/DOF of shared-memory traffic and ~177
// DFMA/DOF; the design below needs ~120 B/DOF and ~57 DFMA/DOF.
//
// Design
//   * n even.  A "work item" is a PAIR of adjacent-i lines of n points; an element has T = n^2/2 items per
//     direction, one item per lane.  n = 8: T = 32, one WARP per element, no block-level barrier at all
//     (__syncwarp only), four independent warps per CTA.  The other sizes pack G elements onto a group of W warps
//     (Shape<N>: n = 10: 3 elements on 5 warps, 150 of 160 lanes; n = 12: 2 on 5; n = 6: 7 on 4) with one named
//     barrier per stage; surplus lanes mirror the last item so that control flow stays uniform.
//   * Every global access is a 128-bit coalesced LDG/STG (a pair of adjacent i); a warp instruction moves 512
//     contiguous bytes.  The six geometric factors (75 % of the traffic) go straight from HBM to registers and
//     never touch shared memory; loads for slab k+kGeoAhead are in flight while slab k is consumed, and a rolling
//     window of `prefetch.global.L2` (kPf slabs ahead, one line per lane, running on into the group's next element)
//     lets those loads find their lines in L2.  (A bulk prefetch of the whole next element was measured and dropped:
//     it is a whole element-time ahead, ~15 us, and the lines do not survive that long in L2.)
//   * D lives in __constant__ memory.  All contractions are fully unrolled, so every DFMA takes its D entry
//     as a uniform-register operand (LDCU): no vector register, no shared-memory read for D.
//   * The three 1-D contractions (and their transposes) are done by the lane that owns the whole line in
//     registers: 2 x n inputs -> 2 x n outputs, n^2 DFMAs each, 2n independent accumulation chains.  Between
//     directions the data is transposed through two or three n^3 shared-memory buffers using only LDS.128/STS.128;
//     the buffer layout (slab stride, padded element stride, XOR swizzles) and the choice of the two rows a lane takes
//     in the i-line stages (RowTable) make all three access patterns bank-conflict-free: tools/check_banks.py models
//     them and reproduces the wavefront counts of the ncu source page.
//   * FP64 FMA on the CUDA cores; tensor cores are not used (the kernel is HBM-bound).
//
// Build: this file is compiled once per supported n with -DNOMPK_AX_N=<n> (the kernels of that n, its own copy of D
// in constant memory and the entry point nompk_ax_run_n<n>; the four units compile in parallel) and once without the
// macro (the C ABI of include/nompk.h, which validates the operands and forwards to the unit of n).
#include <type_traits>
#include <utility>

#include "nompk_common.cuh"
#include "nompk_gridreduce.cuh"

#ifndef NOMPK_AX_N
#define NOMPK_AX_N 0
#endif

namespace nompk {
namespace {

// What the fused p.Ap finish needs (see ax_kernel: kDot); handed from the C ABI unit to the unit of n by pointer.
struct AxDotArgs {
  void *workspace;
  double *result;
  double *result_host;
  unsigned long long host_seq;
  PeerExchange px;  // all-reduce over ranks fused into the finish (world <= 1: none)
};

// p <- r + beta p in front of the operator (see ax_kernel: kXpay).  beta_dev != NULL: beta is read from device memory.
struct AxXpayArgs {
  const double *r = nullptr;
  double *p = nullptr;
  const double *beta_dev = nullptr;
  double beta = 0.0;
};

}  // namespace
}  // namespace nompk

// One entry point per n: stages D (unless NOMPK_AX_D_CACHED), picks the variant, launches.  `dot` is an AxDotArgs or NULL, `xpay` an AxXpayArgs or NULL (needs dot).
#define NOMPK_AX_RUN_DECL(n)                                                                                            \
  extern "C" __attribute__((visibility("hidden"))) int nompk_ax_run_n##n(int variant, size_t E, const double *u,       \
                                                                         const double *g, const double *D, double *w,  \
                                                                         unsigned flags, cudaStream_t stream,          \
                                                                         const void *dot, const void *xpay)
NOMPK_AX_RUN_DECL(6);
NOMPK_AX_RUN_DECL(8);
NOMPK_AX_RUN_DECL(10);
NOMPK_AX_RUN_DECL(12);

// The shapes kept for profiling are compiled in two more units per n (-DNOMPK_AX_PART=1, 2; the production kernels and the
// fused forms are part 0): build time -- one unit with all the kernels of n = 12 takes ten minutes.  Every unit has its
// own __constant__ copy of D, so these entry points stage D themselves.  NOMPK_AX_NO_SUCH_VARIANT: not one of mine.
#ifndef NOMPK_AX_PART
#define NOMPK_AX_PART 0
#endif
#define NOMPK_AX_NO_SUCH_VARIANT (-4711)
#define NOMPK_AX_VARIANTS_DECL(part, n)                                                                                 \
  extern "C" __attribute__((visibility("hidden"))) int nompk_ax_variants##part##_n##n(                                  \
      int variant, size_t E, const double *u, const double *g, const double *D, double *w, cudaStream_t stream)

#if NOMPK_AX_N != 0

namespace nompk {
namespace {

// D[a][l] at nompk_ax_cD[a * n + l] (constant bank 3; read with LDCU into uniform registers, see ld_D); one copy per
// compiled n.
__constant__ double nompk_ax_cD[12 * 12];
// Even-odd split of a centro-antisymmetric D (D[a][l] == -D[n-1-a][n-1-l], what a differentiation matrix on symmetric
// nodes is): four (n/2)^2 tables, [2 * TRANS + odd][a * (n/2) + l] = (M(a,l) +- M(a,n-1-l)) / 2 with M = D or D^T.  See
// eo_apply.  Written by ax_stage_eo when the caller vouches for the symmetry (NOMPK_AX_D_ANTISYMMETRIC).
__constant__ double nompk_ax_cEO[4 * 36];


// ---------------------------------------------------------------------------------------------------
// Shared-memory layout of one n^3 buffer, in 16-byte chunks (one chunk = two adjacent-i doubles).
// ---------------------------------------------------------------------------------------------------
template <int N> struct Layout {
  static constexpr int NP = N / 2;  // chunks per i-line
  static constexpr int SK = (N == 8) ? 36 : (N == 10) ? 53 : (N == 6) ? 19 : (N == 12) ? 78 : N * NP + 1;
  static constexpr int kChunks = N * SK;
  static constexpr int T = N * N / 2;              // work items per element and direction
  static constexpr int WPE = (T + 31) / 32;        // warps per element
  static constexpr int LPE = 32 * WPE;             // lanes per element
  // A 128-bit shared-memory access is served 8 lanes per wavefront; the 8 lanes are conflict-free when their chunk
  // addresses differ mod 8 (8 x 16 B = all 32 banks).  SK = NP (mod 8) makes the k-column pattern (address = tt + const)
  // and the j-line pattern (address = SK q + p = tt (mod 8)) walk through consecutive residues.  Groups of several
  // elements (N != 8) put element el at lanes [el T, (el + 1) T): the progression continues across the element boundary
  // when the elements' buffers are T (mod 8) chunks apart -- kElemStride pads for that.
  template <int BUFS>
  static constexpr int kElemStride = N == 8 ? BUFS * kChunks : BUFS * kChunks + (((T - BUFS * kChunks) % 8) + 8) % 8;
  // N = 12: the rows are 6 chunks long, every row starts at an even chunk and an i-line access (all lanes at the same
  // chunk c of their own row) could only ever reach 4 of the 8 residues.  Bit 2 of k + j flips the chunks of a row in
  // pairs (p ^ 1 stays inside the row), which gives the row a parity; pairs are never split by a wavefront boundary in
  // the other two patterns (8 lanes cut a row of 6 at an even chunk), so those stay conflict-free.
  __device__ static __forceinline__ int swz(int k, int j) {
    if constexpr (N == 8) return (j >> 1) & 3;
    else if constexpr (N == 12) return ((k + j) >> 2) & 1;
    else return 0;
  }
  __device__ static __forceinline__ int at(int k, int j, int p) { return k * SK + j * NP + (p ^ swz(k, j)); }
};

// The i-line stages (S2, S6) give every lane two whole rows (k, j) of an element; which two is free.  With the layouts
// above a row's residue class is (k + j) mod 8 for N = 6, 10, 12 (row base = SK k + NP j = NP (k + j) (mod 8), NP odd; for
// N = 12 the parity swizzle supplies the missing bit), so lane tt takes a row of class tt (mod 8) as its first row and one
// of class tt + 1 as its second: any 8 consecutive lanes then hit 8 different residues, also across the element boundaries
// of a group (see kElemStride).  Adjacent rows (2 tt, 2 tt + 1), the obvious choice, cost 7 (N = 10) and 10 (N = 12)
// wavefronts per access instead of 4 (ncu, profiles/r02_*).  Where the classes do not divide evenly (N = 12: 16 to 20 rows
// per class, 18 needed) the few surplus rows go where they collide least.  Built at compile time; entry 2 tt, 2 tt + 1 = the
// rows (k * N + j) of work item tt.
template <int N> struct RowTable {
  static constexpr int T = N * N / 2;
  unsigned short rows[2 * T];
  constexpr RowTable() : rows() {
    bool used[N * N] = {};
    for (int which = 0; which < 2; which++) {
      for (int tt = 0; tt < T; tt++) {
        const int want = (tt + which) % 8;
        int pick = -1;
        for (int pass = 0; pass < 8 && pick < 0; pass++) {   // the wanted class first, then the nearest ones
          const int cls = (want + pass) % 8;
          for (int r = 0; r < N * N && pick < 0; r++)
            if (!used[r] && (r / N + r % N) % 8 == cls) pick = r;
        }
        used[pick] = true;
        rows[2 * tt + which] = (unsigned short)pick;
      }
    }
  }
};

// (N = 8 keeps the adjacent rows 2 tt, 2 tt + 1: its XOR swizzle already makes them conflict-free.)
template <int N> __constant__ RowTable<N> nompk_ax_rows = RowTable<N>();

__device__ __forceinline__ double2 ldg2(const double2 *p) { return __ldg(p); }

// A weak (coherent) load: unlike an .nc load, ptxas may not sink it below a barrier (see kPinGeo in ax_kernel).
__device__ __forceinline__ double2 ldg2_ordered(const double2 *p) {
  double2 r;
  asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

__device__ __forceinline__ double2 ldg2_stream(const double2 *p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

__device__ __forceinline__ void prefetch_l2_bulk(const void *p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// the same, asking L2 to keep the line until somebody uses it (the streams of w and of the other CTAs' factors push a
// line that was prefetched with normal priority out again before its demand load arrives: 14 - 30 % of the factors are
// read from DRAM twice with three CTAs per SM, profiles/r02_ncu_summary.md, section 4)
__device__ __forceinline__ void prefetch_l2_keep(const void *p) {
  asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p));
}

// ... and the demand load that uses such a line tells L2 it is the first to go
__device__ __forceinline__ double2 ldg2_last_use(const double2 *p) {
  double2 r;
  unsigned long long policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(policy));
  return r;
}

// Groups whose lanes outnumber their work items (n = 10: 160 lanes, 150 items) let the surplus lanes mirror the last
// item: same loads, same arithmetic, same stores of the same values to the same addresses -- control flow stays
// uniform (see ax_kernel).  Where a stage reads and then overwrites one shared-memory location (S4, S7) the mirrors
// and their original sit in the same warp; this fence orders all the reads of the warp before its writes without
// relying on lock-step execution.  Compiles to nothing for groups without surplus lanes.
template <bool kHasMirrors> __device__ __forceinline__ void mirror_fence() {
  if constexpr (kHasMirrors) __syncwarp();
}

template <int GL> __device__ __forceinline__ void element_sync(int grp) {
  if constexpr (GL == 32) {
    __syncwarp();
  } else {
    asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(GL) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// D entries as instruction operands.
// In uniform control flow ptxas serves a __constant__ read with LDCU (uniform datapath) and feeds the DFMA from
// the uniform register: a D entry costs no vector register and no LSU slot.  Left alone, though, both NVVM and
// ptxas treat the n^2 reads as loop invariants, hoist them out of the element loop and pin 2 n^2 VECTOR registers
// for the whole kernel (R2UR before every DFMA, heavy spilling).  So every stage indexes D with its own
// "zero" that is derived from the loop counter (iteration counter >> 24..29: always 0, but not provably so): the reads become
// loop-variant, stay next to their DFMAs as `LDCU.64 URx, c[3][URz + imm]`, and are not merged across stages.
// ---------------------------------------------------------------------------------------------------
template <int IDX> __device__ __forceinline__ double ld_D(int z) { return nompk_ax_cD[IDX + z]; }

template <int... Is, typename F>
__device__ __forceinline__ void static_for(std::integer_sequence<int, Is...>, F &&f) {
  (f(std::integral_constant<int, Is>{}), ...);
}

// M(a, l) = D[a][l] (TRANS = false) or D[l][a] (TRANS = true)
template <int N, bool TRANS, int A, int L_> struct DIdx {
  static constexpr int value = TRANS ? L_ * N + A : A * N + L_;
};

// s += sum_l M(A, l) * in[l], for a PAIR of lines packed as double2 (same D entry for both).
template <int N, bool TRANS, int A>
__device__ __forceinline__ double2 dot_pair(const double2 (&in)[N], double2 s, int z) {
  static_for(std::make_integer_sequence<int, N>{}, [&](auto Lc) {
    constexpr int l = decltype(Lc)::value;
    const double d = ld_D<DIdx<N, TRANS, A, l>::value>(z);
    s.x = fma(d, in[l].x, s.x);
    s.y = fma(d, in[l].y, s.y);
  });
  return s;
}

// Two i-lines held as chunks: v0[c], v1[c] = points (2c, 2c+1) of row 0 / row 1.  Returns output points
// (2C, 2C+1) of both rows.
template <int N, bool TRANS, int C>
__device__ __forceinline__ void dot_rows(const double2 (&v0)[N / 2], const double2 (&v1)[N / 2], double2 &o0,
                                         double2 &o1, int z) {
  double2 a0 = make_double2(0.0, 0.0), a1 = make_double2(0.0, 0.0);
  static_for(std::make_integer_sequence<int, N / 2>{}, [&](auto Cc) {
    constexpr int c = decltype(Cc)::value;
    {
      const double d0 = ld_D<DIdx<N, TRANS, 2 * C, 2 * c>::value>(z);
      const double d1 = ld_D<DIdx<N, TRANS, 2 * C + 1, 2 * c>::value>(z);
      a0.x = fma(d0, v0[c].x, a0.x);
      a1.x = fma(d0, v1[c].x, a1.x);
      a0.y = fma(d1, v0[c].x, a0.y);
      a1.y = fma(d1, v1[c].x, a1.y);
    }
    {
      const double d0 = ld_D<DIdx<N, TRANS, 2 * C, 2 * c + 1>::value>(z);
      const double d1 = ld_D<DIdx<N, TRANS, 2 * C + 1, 2 * c + 1>::value>(z);
      a0.x = fma(d0, v0[c].y, a0.x);
      a1.x = fma(d0, v1[c].y, a1.x);
      a0.y = fma(d1, v0[c].y, a0.y);
      a1.y = fma(d1, v1[c].y, a1.y);
    }
  });
  o0 = a0;
  o1 = a1;
}

// ---------------------------------------------------------------------------------------------------
// Even-odd contraction (kEO: a mask of the stages that use it, see kEOAll).  With e[l] = in[l] + in[n-1-l], o[l] = in[l] - in[n-1-l] (l < n/2) and the tables above,
//   out[a]     = sum_l De(a,l) e[l] + Do(a,l) o[l] = P + Q         (any D)
//   out[n-1-a] = Q - P                                             (centro-antisymmetric D)
// n^2/2 + 2n operations per line instead of n^2: 30 % fewer FP64 instructions at n = 10, a third fewer at n = 12.  The
// sums are taken in another order than in dot_pair / dot_rows (and D enters as its symmetrised halves), so the result
// differs from the general path in the last bits -- the caller opts in, bit-exact data keeps the general path.
// ---------------------------------------------------------------------------------------------------
template <int IDX> __device__ __forceinline__ double ld_EO(int z) { return nompk_ax_cEO[IDX + z]; }

template <int N, bool TRANS>
__device__ __forceinline__ void eo_apply(const double2 (&in)[N], double2 (&out)[N], int z) {
  constexpr int H = N / 2;
  double2 e[H], o[H];
#pragma unroll
  for (int l = 0; l < H; l++) {
    e[l].x = in[l].x + in[N - 1 - l].x, e[l].y = in[l].y + in[N - 1 - l].y;
    o[l].x = in[l].x - in[N - 1 - l].x, o[l].y = in[l].y - in[N - 1 - l].y;
  }
  static_for(std::make_integer_sequence<int, H>{}, [&](auto Ac) {
    constexpr int a = decltype(Ac)::value;
    double2 P = make_double2(0.0, 0.0), Q = make_double2(0.0, 0.0);
    static_for(std::make_integer_sequence<int, H>{}, [&](auto Lc) {
      constexpr int l = decltype(Lc)::value;
      const double de = ld_EO<(TRANS ? 72 : 0) + a * H + l>(z);
      const double dq = ld_EO<(TRANS ? 108 : 36) + a * H + l>(z);
      P.x = fma(de, e[l].x, P.x), P.y = fma(de, e[l].y, P.y);
      Q.x = fma(dq, o[l].x, Q.x), Q.y = fma(dq, o[l].y, Q.y);
    });
    out[a].x = P.x + Q.x, out[a].y = P.y + Q.y;
    out[N - 1 - a].x = Q.x - P.x, out[N - 1 - a].y = Q.y - P.y;
  });
}

// The same for two i-lines held as chunks (v[c] = points 2c, 2c + 1): regrouped as pairs (row 0, row 1) per point.
template <int N, bool TRANS>
__device__ __forceinline__ void eo_apply_rows(const double2 (&v0)[N / 2], const double2 (&v1)[N / 2], double2 (&o0)[N / 2],
                                              double2 (&o1)[N / 2], int z) {
  double2 in[N], out[N];
#pragma unroll
  for (int c = 0; c < N / 2; c++) {
    in[2 * c] = make_double2(v0[c].x, v1[c].x);
    in[2 * c + 1] = make_double2(v0[c].y, v1[c].y);
  }
  eo_apply<N, TRANS>(in, out, z);
#pragma unroll
  for (int c = 0; c < N / 2; c++) {
    o0[c] = make_double2(out[2 * c].x, out[2 * c + 1].x);
    o1[c] = make_double2(out[2 * c].y, out[2 * c + 1].y);
  }
}

// One thread per table entry; `eo` is the address of nompk_ax_cEO of the unit that launches it.
__global__ void ax_stage_eo(const double *__restrict__ D, double *__restrict__ eo, int n) {
  const int h = n / 2, idx = threadIdx.x;
  if (idx >= 4 * h * h) return;
  const int t = idx / (h * h), a = (idx / h) % h, l = idx % h;
  const bool trans = t >= 2, odd = t & 1;
  const double m1 = trans ? D[l * n + a] : D[a * n + l];
  const double m2 = trans ? D[(n - 1 - l) * n + a] : D[a * n + (n - 1 - l)];
  eo[t * 36 + a * h + l] = 0.5 * (odd ? m1 - m2 : m1 + m2);
}

// kGeoAhead: how many k-slabs of geometric factors are in flight in registers ahead of their use.
// kPf: L2 prefetch policy.  0 = none; 1 = one bulk prefetch of the whole next element (measured: doubles DRAM reads,
// the 66 MB that 2368 warps keep "reserved" do not survive in L2); >= 2 = rolling window: while slab k of the
// geometric factors is consumed, the 24 lines of slab k + kPf (wrapping into the next element of this warp) are
// requested with one fire-and-forget `prefetch.global.L2` per lane, so demand loads find their lines in L2 and the
// HBM latency is paid by requests that hold no register and no scoreboard slot.
//
// Control flow is kept CTA-uniform on purpose: the element loop depends only on blockIdx; lanes beyond the T work
// items of an element (n = 10: 14 of 64) and elements beyond E in the last group do not branch, they mirror
// the last valid item / element and write the same values to the same addresses.  (Divergent regions would make
// ptxas drop the uniform-register D operands and issue one vector LDC.64 per DFMA.)
//
// kDot: also return u . (A u), the p.Ap of a CG iteration, without reading u or w again: with ur, us, ut the reference-
// space gradient of u and (wr, ws, wt) = G (ur, us, ut), u . A u = sum over points of ur wr + us ws + ut wt, and all six
// values are in registers in the geometric stage.  Finished like libnompk's reductions: block tree, one partial per
// CTA, atomic ticket, the last CTA folds the partials in CTA order (deterministic) and publishes the scalar.
// kPersistent: the grid has as many CTAs as fit on the device and each strides over the elements; otherwise one CTA
// per GPC*G elements (the hardware scheduler streams the CTAs) and `pf_stride` = elements of all resident CTAs is only
// the distance of the L2 prefetch (what will be scheduled next on SOME SM).
//
// kTwoBuf (experimental, not the default anywhere yet): two shared buffers per element instead of three.  us is computed
// first (B0 -> B2); after a barrier ur replaces u IN PLACE in B0 (a lane reads its two i-lines whole before it writes
// them), wr then replaces ur in place and D_r^T wr replaces wr in place.  One barrier more per element, a third less
// shared memory: at n = 10 / 12 a third CTA fits an SM once the register budget allows it.
//
// kXpay: the direction update of conjugate gradients, p <- r + beta p, in front of the operator: the lane that loads its
// k-column pair of p for S0 loads the same pair of r, forms the new p in registers (multiply, then add: the roundings of
// the stand-alone map kernel), stores it and carries on with it as u.  One pass over p and r less per CG iteration
// (80 instead of 24 + 64 B/DOF) and one launch less.  p is read and written through the same non-const pointer by the
// one lane that owns the pair; nobody else touches it.
// (The xpay arguments are a trailing kernel parameter that is an empty structure for every other instantiation, so the
// kernels without it keep their parameter layout -- and their SASS.)
struct AxNoXpay {};

//
// kPfMode: shape and eviction priority of the rolling L2 prefetch.
//   0: the window runs on into the group's NEXT element (its first kPf slabs and its u are requested during the last
//      slabs of this one), normal priority;  1: the same, prefetched lines are kept (evict_last);  2: kept, and the demand
//      load of a factor marks its line as the first to go (evict_first).
//   3: LOCAL window -- every element warms its own window in S0 (slabs kGeoAhead .. kPf - 1, consumed after S1 - S3) and
//      S4 only prefetches inside the element; the next element's u is requested after S6.  Every prefetch then leads its
//      demand load by 2 - 4 us.  With modes 0 - 2 the lines of the next element wait for a whole S5 .. S8, S0 .. S3 --
//      about 9 us with three CTAs per SM -- and L2 (126 MB under 8.5 TB/s of traffic: ~12 us of residence) has dropped
//      part of them by then: 14 % (n = 10) to 30 % (n = 12) of the factors came from DRAM twice (profiles/r02_ncu_summary.md, section 4).
//   4: local window, and demand loads mark their lines evict_first.
//   5: local window, and the first kGeoAhead slabs are prefetched too, at the start of the element, instead of being
//      fetched by the loads that fill the register ring (a mode 6 did it at the end of the previous element).  Measured:
//      no gain (profiles/r02_dot_pin_experiment.jsonl) -- the stall in front of S4 is L2 latency, not DRAM latency.
// kPin: the loads that fill the register ring stand in the source at S0 ("in flight during S1 .. S3"), but at 128
//   registers ptxas sinks them to the end of the stage before S4 -- and in the kernels with the fused dot product, whose
//   two more live values sit on the register cliff of S4, BELOW the barrier in front of S4, next to their use: the fused
//   kernel was 21 % slower than the plain one at n = 10 (long-scoreboard stalls per issue 5.5 instead of 3.2).  kPin = 1
//   issues them at the end of the stage before S4 as weak loads (ld.global, not .nc), which ptxas may not move across the
//   barrier: 0.71 -> 0.79 of the peak at n = 10 (the plain kernel: 0.87 on the same box).  No gain at n = 6, 12 or for
//   the plain kernels.  Also tried for the fused kernel and dropped: u . w formed in S8 from a copy of u in shared
//   memory instead of the energy form in S4 (slower: 0.71 - 0.74), 32-bit element numbers (more spills).
template <int N, int G, int W, int GPC, int kGeoAhead, int kPf, bool kStreamLoads, int kMinBlocks, bool kDot,
          bool kPersistent, bool kTwoBuf = false, bool kXpay = false, int kPfMode = 0, int kPin = 0, int kEO = 0>
__global__ void __launch_bounds__(GPC * W * 32, kMinBlocks)
ax_kernel(const double *__restrict__ u, const double *__restrict__ g, double *__restrict__ w, size_t E, AxDotArgs dot,
          size_t pf_stride, std::conditional_t<kXpay, AxXpayArgs, AxNoXpay> xp) {
  using L = Layout<N>;
  using SeqN = std::make_integer_sequence<int, N>;
  using SeqNP = std::make_integer_sequence<int, N / 2>;
  // A GROUP of W warps works on G elements at a time: G*T work items on 32*W lanes (n = 8: 1 element per warp,
  // 100 %; n = 10: 3 elements on 5 warps, 150/160 lanes = 94 % instead of 50/64 = 78 % with one element on 2 warps).
  constexpr int NP = L::NP, T = L::T, GL = 32 * W;
  static_assert(G * T <= GL, "group too small for its elements");
  constexpr int N3 = N * N * N;
  constexpr int SLAB2 = N * NP;  // double2 per k-slab in global memory
  extern __shared__ double2 smem[];

  const int grp = threadIdx.x / GL;             // group within the CTA
  const int gid = threadIdx.x % GL;             // lane within the group
  const int gidc = gid < G * T ? gid : G * T - 1;  // surplus lanes mirror the last work item (no divergence)
  const int el = gidc / T;                      // element within the group
  const int tt = gidc % T;                      // work item within the element
  const int t = tt;
  // k-column item: (p, j);  j-line item: (p, k) -- the same split of t.
  const int p = tt % NP, q = tt / NP;
  // i-line items: two rows (k, j) of the element per lane -- rows 2t and 2t + 1 for N = 8, the conflict-free assignment
  // of RowTable otherwise.  oa / ob: chunk offset of the row, sa / sb: its swizzle (chunk c sits at o + (c ^ s)).
  const int row_a = N == 8 ? 2 * tt : nompk_ax_rows<N>.rows[2 * tt], row_b = N == 8 ? 2 * tt + 1 : nompk_ax_rows<N>.rows[2 * tt + 1];
  const int oa = (row_a / N) * L::SK + (row_a % N) * NP, sa = L::swz(row_a / N, row_a % N);
  const int ob = (row_b / N) * L::SK + (row_b % N) * NP, sb = L::swz(row_b / N, row_b % N);

  constexpr int kBufs = kTwoBuf ? 2 : 3;
  double2 *B0 = smem + (size_t)(grp * G + el) * L::template kElemStride<kBufs>;
  double2 *B1 = kTwoBuf ? B0 : B0 + L::kChunks;   // ur, then wr
  double2 *B2 = B0 + (kBufs - 1) * L::kChunks;    // us, then ws

  constexpr int EPB = GPC * G;  // elements per CTA and iteration
  const size_t estride = kPersistent ? (size_t)gridDim.x * EPB : pf_stride;
  int it = 0;
  double energy = 0.0;  // kDot: this lane's share of u . A u
  for (size_t eb = (size_t)blockIdx.x * EPB; eb < E; eb += (kPersistent ? estride : E)) {
    // always-zero, loop-variant offsets for the D reads (see ld_D): one per stage, derived from a counter that
    // depends on nothing but the iteration number so that ptxas keeps it in a uniform register
    const int z1 = it >> 24, z2 = it >> 25, z3 = it >> 26, z5 = it >> 27, z6 = it >> 28, z7 = it >> 29;
    it++;
    size_t e = eb + grp * G + el;
    // lanes that only mirror another work item / element must not contribute to the dot product
    const bool dot_counts = gid < G * T && e < E;
    const bool e_valid = e < E;  // kXpay: an element that only mirrors the last one must not store (see S0)
    e = e < E ? e : E - 1;
    const double2 *ue = reinterpret_cast<const double2 *>(u + e * N3) + q * NP + p;
    const double2 *ge = reinterpret_cast<const double2 *>(g + e * 6 * N3) + q * NP + p;
    double2 *we = reinterpret_cast<double2 *>(w + e * N3) + q * NP + p;

    if constexpr (kPf == 1) {
      const size_t en = e + estride;
      if (t == 0 && en < E) {
        prefetch_l2_bulk(u + en * N3, N3 * 8);
        prefetch_l2_bulk(g + en * 6 * N3, 6 * N3 * 8);
      }
    }
    // Rolling prefetch: line `pl` of factor `pf` of a slab (a slab of one factor is N*N*8 bytes).
    constexpr int kLinesPerFactor = (N * N * 8 + 127) / 128;
    const int pf_f = t / kLinesPerFactor, pf_l = t % kLinesPerFactor;
    const size_t e_next = (e + estride < E) ? e + estride : e;
    constexpr bool kLocalWindow = kPfMode >= 3;
    auto prefetch_slab = [&](size_t ee, int ks) {
      const double *line = g + (ee * 6 + pf_f) * N3 + ks * N * N + pf_l * 16;
      if (pf_f < 6) {
        if constexpr (kPfMode == 1 || kPfMode == 2) prefetch_l2_keep(line);
        else prefetch_l2(line);
      }
    };
    auto load_g = [&](const double2 *src) {
      if constexpr (kPfMode == 2 || kPfMode == 4) return ldg2_last_use(src);
      else return kStreamLoads ? ldg2_stream(src) : ldg2(src);
    };
    auto prefetch_next_u = [&]() {
      constexpr int kULines = (N3 * 8 + 127) / 128;
#pragma unroll
      for (int l0 = 0; l0 < kULines; l0 += T)
        if (l0 + t < kULines) prefetch_l2(u + e_next * N3 + (l0 + t) * 16);
    };
    if constexpr (kPf >= 2) {
      if (kLocalWindow || eb == (size_t)blockIdx.x * EPB) {  // warm the window (wrapping window: first element of the CTA only)
        // the first kGeoAhead slabs are fetched by the loads that fill the register ring -- which ptxas issues only in
        // the stage before S4 (register pressure).  Mode 5 prefetches them here as well (measured: no gain).
        constexpr int kFirst = kPfMode == 5 ? 0 : kGeoAhead;
#pragma unroll
        for (int k = kFirst; k < kPf && k < N; k++) prefetch_slab(e, k);
      }
    }

    // ---- S0: u k-column pair -> registers and B0 -----------------------------------------------------
    double2 col[N];  // u column, later ut, later wt
    double2 gq[kGeoAhead][6];
    if constexpr (kXpay) {
      const double beta = xp.beta_dev ? xp.beta_dev[0] : xp.beta;
      const double2 *re = reinterpret_cast<const double2 *>(xp.r + e * N3) + q * NP + p;
      double2 *pe = reinterpret_cast<double2 *>(xp.p + e * N3) + q * NP + p;
      // The update is in place and therefore not idempotent, unlike everything else mirrored lanes repeat.  Surplus
      // lanes of a group sit in the warp of the lane they mirror: all of them have read the old p before anybody stores
      // the new one (fence), and then store the same values.  A whole ELEMENT that mirrors the last one (partial last
      // group) may read a half-updated p: it stores nothing, here and in S8, and its dot weight is zero.
#pragma unroll
      for (int k = 0; k < N; k++) col[k] = pe[k * SLAB2];
#pragma unroll
      for (int k = 0; k < N; k++) {
        const double2 rk = ldg2(re + k * SLAB2);
        col[k].x = __dadd_rn(rk.x, __dmul_rn(beta, col[k].x));
        col[k].y = __dadd_rn(rk.y, __dmul_rn(beta, col[k].y));
      }
      mirror_fence<G * T < GL>();
#pragma unroll
      for (int k = 0; k < N; k++)
        if (e_valid) pe[k * SLAB2] = col[k];
    } else {
#pragma unroll
      for (int k = 0; k < N; k++) col[k] = kStreamLoads ? ldg2_stream(ue + k * SLAB2) : ldg2(ue + k * SLAB2);
    }
    // first slabs of geometric factors: in flight during S1..S3 (as far as the registers allow: at 128 registers ptxas
    // issues them in the last stage before S4).  kPinGeo: the dot product costs the registers that made ptxas sink most
    // of these loads BELOW the barrier in front of S4, next to their use (n = 10: 0.71 instead of 0.87 of the peak);
    // there they are issued in the stage before S4 as weak loads, which may not cross the barrier.
    constexpr bool kPinGeo = kPin != 0 && kTwoBuf;
    if constexpr (!kPinGeo) {
#pragma unroll
      for (int a = 0; a < kGeoAhead; a++)
#pragma unroll
        for (int f = 0; f < 6; f++)
          gq[a][f] = load_g(ge + f * (N3 / 2) + a * SLAB2);
    }
#pragma unroll
    for (int k = 0; k < N; k++) B0[L::at(k, q, p)] = col[k];
    // ---- S1: ut = D_t u along the k-column, in registers ---------------------------------------------
    {
      double2 ut[N];
      if constexpr ((kEO & 1) != 0) {
        eo_apply<N, false>(col, ut, z1);
      } else {
        static_for(SeqN{}, [&](auto A) {
          constexpr int a = decltype(A)::value;
          ut[a] = dot_pair<N, false, a>(col, make_double2(0.0, 0.0), z1);
        });
      }
#pragma unroll
      for (int k = 0; k < N; k++) col[k] = ut[k];
    }
    element_sync<GL>(grp);

    if constexpr (kTwoBuf) {
      // ---- S3 first: us = D_s u on a j-line pair (p, k = q): B0 -> B2 -----------------------------------
      {
        double2 in[N];
#pragma unroll
        for (int l = 0; l < N; l++) in[l] = B0[L::at(q, l, p)];
        if constexpr ((kEO & 2) != 0) {
          double2 us[N];
          eo_apply<N, false>(in, us, z3);
#pragma unroll
          for (int j = 0; j < N; j++) B2[L::at(q, j, p)] = us[j];
        } else {
          static_for(SeqN{}, [&](auto A) {
            constexpr int j = decltype(A)::value;
            B2[L::at(q, j, p)] = dot_pair<N, false, j>(in, make_double2(0.0, 0.0), z3);
          });
        }
      }
      element_sync<GL>(grp);  // nobody reads u from B0 any more
      // ---- S2: ur = D_r u on two i-lines, in place in B0 ---------------------------------------------------
      {
        double2 v0[NP], v1[NP];
#pragma unroll
        for (int c = 0; c < NP; c++) v0[c] = B0[oa + (c ^ sa)], v1[c] = B0[ob + (c ^ sb)];
        mirror_fence<G * T < GL>();  // the mirrors of a lane have read the same two lines
        if constexpr ((kEO & 4) != 0) {
          double2 r0[NP], r1[NP];
          eo_apply_rows<N, false>(v0, v1, r0, r1, z2);
#pragma unroll
          for (int c = 0; c < NP; c++) B0[oa + (c ^ sa)] = r0[c], B0[ob + (c ^ sb)] = r1[c];
        } else {
          static_for(SeqNP{}, [&](auto C) {
            constexpr int c = decltype(C)::value;
            double2 o0, o1;
            dot_rows<N, false, c>(v0, v1, o0, o1, z2);
            B0[oa + (c ^ sa)] = o0;
            B0[ob + (c ^ sb)] = o1;
          });
        }
      }
      if constexpr (kPinGeo) {
#pragma unroll
        for (int a = 0; a < kGeoAhead; a++)
#pragma unroll
          for (int f = 0; f < 6; f++)
            gq[a][f] = ldg2_ordered(ge + f * (N3 / 2) + a * SLAB2);
      }
      element_sync<GL>(grp);
    } else {
      // ---- S2: ur = D_r u on two i-lines -> B1 ----------------------------------------------------------
      {
        double2 v0[NP], v1[NP];
  #pragma unroll
        for (int c = 0; c < NP; c++) v0[c] = B0[oa + (c ^ sa)], v1[c] = B0[ob + (c ^ sb)];
        if constexpr ((kEO & 4) != 0) {
          double2 r0[NP], r1[NP];
          eo_apply_rows<N, false>(v0, v1, r0, r1, z2);
#pragma unroll
          for (int c = 0; c < NP; c++) B1[oa + (c ^ sa)] = r0[c], B1[ob + (c ^ sb)] = r1[c];
        } else {
          static_for(SeqNP{}, [&](auto C) {
            constexpr int c = decltype(C)::value;
            double2 o0, o1;
            dot_rows<N, false, c>(v0, v1, o0, o1, z2);
            B1[oa + (c ^ sa)] = o0;
            B1[ob + (c ^ sb)] = o1;
          });
        }
      }
      // ---- S3: us = D_s u on a j-line pair (p, k = q) -> B2 ---------------------------------------------
      {
        double2 in[N];
  #pragma unroll
        for (int l = 0; l < N; l++) in[l] = B0[L::at(q, l, p)];
        if constexpr ((kEO & 2) != 0) {
          double2 us[N];
          eo_apply<N, false>(in, us, z3);
#pragma unroll
          for (int j = 0; j < N; j++) B2[L::at(q, j, p)] = us[j];
        } else {
          static_for(SeqN{}, [&](auto A) {
            constexpr int j = decltype(A)::value;
            B2[L::at(q, j, p)] = dot_pair<N, false, j>(in, make_double2(0.0, 0.0), z3);
          });
        }
      }
      element_sync<GL>(grp);
    }

    // ---- S4: geometric factors at the k-column (p, j = q); wr -> B1, ws -> B2, wt -> registers ---------
#pragma unroll
    for (int k = 0; k < N; k++) {
      const int slot = k % kGeoAhead;
      double2 gk[6];
#pragma unroll
      for (int f = 0; f < 6; f++) gk[f] = gq[slot][f];
      if (k + kGeoAhead < N) {
#pragma unroll
        for (int f = 0; f < 6; f++)
          gq[slot][f] = load_g(ge + f * (N3 / 2) + (k + kGeoAhead) * SLAB2);
      }
      if constexpr (kPf >= 2) {
        // slab k + kPf of this element, or slab (k + kPf - N) of the next one; next element's u with slab 0
        const int ks = k + kPf;
        if (ks < N) prefetch_slab(e, ks);
        else if (!kLocalWindow && ks - N < N) prefetch_slab(e_next, ks - N);
        if (!kLocalWindow && k == 0) prefetch_next_u();
      }
      const int a = L::at(k, q, p);
      const double2 ur = B1[a], us = B2[a], ut = col[k];
      double2 wr, ws, wt;
      wr.x = fma(gk[0].x, ur.x, fma(gk[1].x, us.x, gk[2].x * ut.x));
      wr.y = fma(gk[0].y, ur.y, fma(gk[1].y, us.y, gk[2].y * ut.y));
      ws.x = fma(gk[1].x, ur.x, fma(gk[3].x, us.x, gk[4].x * ut.x));
      ws.y = fma(gk[1].y, ur.y, fma(gk[3].y, us.y, gk[4].y * ut.y));
      wt.x = fma(gk[2].x, ur.x, fma(gk[4].x, us.x, gk[5].x * ut.x));
      wt.y = fma(gk[2].y, ur.y, fma(gk[4].y, us.y, gk[5].y * ut.y));
      if constexpr (kDot) {
        double ex = fma(ur.x, wr.x, fma(us.x, ws.x, ut.x * wt.x));
        double ey = fma(ur.y, wr.y, fma(us.y, ws.y, ut.y * wt.y));
        // select, not multiply by a weight: with kXpay an element that mirrors the last one may have read a half-updated
        // p -- its numbers are finite but meaningless, and 0 * inf would let an overflow through; and a predicate costs
        // no register pair (at 128 registers the weight pushed the first loads of the geometric factors behind the
        // barrier of S4: 0.71 instead of 0.87 of the peak at n = 10)
        energy += dot_counts ? ex + ey : 0.0;
      }
      mirror_fence<G * T < GL>();
      B1[a] = wr;
      B2[a] = ws;
      col[k] = wt;
    }
    // ---- S5: w = D_t^T wt along the k-column, in registers --------------------------------------------
    double2 wacc[N];
    if constexpr ((kEO & 8) != 0) {
      eo_apply<N, true>(col, wacc, z5);
    } else {
      static_for(SeqN{}, [&](auto A) {
        constexpr int a = decltype(A)::value;
        wacc[a] = dot_pair<N, true, a>(col, make_double2(0.0, 0.0), z5);
      });
    }
    element_sync<GL>(grp);

    // ---- S6: D_r^T wr on two i-lines: B1 -> B0 ----------------------------------------------------------
    {
      double2 v0[NP], v1[NP];
#pragma unroll
      for (int c = 0; c < NP; c++) v0[c] = B1[oa + (c ^ sa)], v1[c] = B1[ob + (c ^ sb)];
      if constexpr (kTwoBuf) mirror_fence<G * T < GL>();  // B1 is B0: in place, mirrors must have read first
      if constexpr ((kEO & 16) != 0) {
        double2 r0[NP], r1[NP];
        eo_apply_rows<N, true>(v0, v1, r0, r1, z6);
#pragma unroll
        for (int c = 0; c < NP; c++) B0[oa + (c ^ sa)] = r0[c], B0[ob + (c ^ sb)] = r1[c];
      } else {
        static_for(SeqNP{}, [&](auto C) {
          constexpr int c = decltype(C)::value;
          double2 o0, o1;
          dot_rows<N, true, c>(v0, v1, o0, o1, z6);
          B0[oa + (c ^ sa)] = o0;
          B0[ob + (c ^ sb)] = o1;
        });
      }
    }
    element_sync<GL>(grp);
    if constexpr (kPf >= 2 && kLocalWindow) prefetch_next_u();

    // ---- S7: + D_s^T ws on the j-line pair (p, k = q): B2, B0 -> B0 -------------------------------------
    {
      double2 in[N];
#pragma unroll
      for (int l = 0; l < N; l++) in[l] = B2[L::at(q, l, p)];
      if constexpr ((kEO & 32) != 0) {
        double2 ws[N];
        eo_apply<N, true>(in, ws, z7);
#pragma unroll
        for (int j = 0; j < N; j++) {
          const double2 r = B0[L::at(q, j, p)];
          ws[j].x += r.x, ws[j].y += r.y;
        }
        mirror_fence<G * T < GL>();
#pragma unroll
        for (int j = 0; j < N; j++) B0[L::at(q, j, p)] = ws[j];
      } else {
        static_for(SeqN{}, [&](auto A) {
          constexpr int j = decltype(A)::value;
          const int a = L::at(q, j, p);
          const double2 sum = dot_pair<N, true, j>(in, B0[a], z7);
          mirror_fence<G * T < GL>();
          B0[a] = sum;
        });
      }
    }
    element_sync<GL>(grp);

    // ---- S8: add the r/s part to the k-column and store w ------------------------------------------------
#pragma unroll
    for (int k = 0; k < N; k++) {
      const double2 rs = B0[L::at(k, q, p)];
      double2 o;
      o.x = wacc[k].x + rs.x;
      o.y = wacc[k].y + rs.y;
      if constexpr (kXpay) {
        if (e_valid) we[k * SLAB2] = o;
      } else {
        we[k * SLAB2] = o;
      }
    }
    // No group barrier needed here: the next iteration's S0 writes exactly the B0 chunks this lane just read.  Its
    // mirrors read the same chunks, though, and must have done so before anybody overwrites them.
    mirror_fence<G * T < GL>();
  }

  if constexpr (kDot) {
    constexpr int kWarps = GPC * W;
    __shared__ double warp_part[kWarps];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) energy += __shfl_xor_sync(0xffffffffu, energy, off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_part[warp] = energy;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
      for (int i = 0; i < kWarps; i++) s += warp_part[i];
    struct Sum {
      static __device__ __forceinline__ double identity() { return 0.0; }
      static __device__ __forceinline__ double combine(double a, double b) { return a + b; }
    };
    grid_finish<Sum, double, (kWarps * 32 < 64 ? 64 : kWarps * 32)>(s, dot.workspace, dot.result, dot.result_host, dot.host_seq,
                                                                    dot.px);
  }
}

template <int N, int G, int W, int GPC, int GA, int PF, bool ST, int MB, bool DOT = false, bool PERSISTENT = true,
          bool TWOBUF = false, int PFMODE = 0, int PIN = 0, int EO = 0>
int launch_ax(size_t E, const double *u, const double *g, double *w, cudaStream_t stream, AxDotArgs dot = AxDotArgs()) {
  using L = Layout<N>;
  auto kern = ax_kernel<N, G, W, GPC, GA, PF, ST, MB, DOT, PERSISTENT, TWOBUF, false, PFMODE, PIN, EO>;
  constexpr int kThreads = GPC * W * 32, kElems = GPC * G;
  const size_t smem = (size_t)kElems * L::template kElemStride<(TWOBUF ? 2 : 3)> * sizeof(double2);
  static bool configured = false;
  static int blocks_per_sm = 1;
  if (!configured) {
    NOMPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NOMPK_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, kThreads, smem));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    configured = true;
  }
  size_t blocks = (E + kElems - 1) / kElems;
  const size_t cap = (size_t)sm_count() * blocks_per_sm;
  if (PERSISTENT && blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, kThreads, smem, stream>>>(u, g, w, E, dot, cap * kElems, AxNoXpay());
  NOMPK_LAUNCH_CHECK("ax_kernel");
  return NOMPK_OK;
}

template <int N, int G, int W, int GPC, int GA, int PF, int MB, int EO = 0>
int launch_ax_xpay_dot(size_t E, const double *g, double *w, cudaStream_t stream, AxDotArgs dot, AxXpayArgs xp) {
  using L = Layout<N>;
  auto kern = ax_kernel<N, G, W, GPC, GA, PF, false, MB, true, true, false, true, (N == 8 ? 0 : 3), 0, EO>;
  constexpr int kThreads = GPC * W * 32, kElems = GPC * G;
  const size_t smem = (size_t)kElems * L::template kElemStride<3> * sizeof(double2);
  static bool configured = false;
  static int blocks_per_sm = 1;
  if (!configured) {
    NOMPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NOMPK_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, kThreads, smem));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    configured = true;
  }
  size_t blocks = (E + kElems - 1) / kElems;
  const size_t cap = (size_t)sm_count() * blocks_per_sm;
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, kThreads, smem, stream>>>(xp.p, g, w, E, dot, cap * kElems, xp);
  NOMPK_LAUNCH_CHECK("ax_kernel (xpay)");
  return NOMPK_OK;
}

// kEO is a mask of the stages that take the even-odd form: 1 S1, 2 S3, 4 S2, 8 S5, 16 S6, 32 S7.
constexpr int kEOAll = 63;

// Group shapes: <elements per group, warps per group, groups per CTA>.
template <int N> struct Shape;
template <> struct Shape<6> { static constexpr int G = 7, W = 4, GPC = 1; };    // 126 / 128 lanes
template <> struct Shape<8> { static constexpr int G = 1, W = 1, GPC = 4; };    // 32 / 32, warps independent
template <> struct Shape<10> { static constexpr int G = 3, W = 5, GPC = 1; };   // 150 / 160
template <> struct Shape<12> { static constexpr int G = 2, W = 5, GPC = 1; };   // 144 / 160

#if NOMPK_AX_PART == 0
// Ax fused with u . A u (production shapes only).
// Even-odd contractions where they pay (interleaved sweeps, profiles/r02_even_odd.jsonl): n = 8 and n = 10 in every
// stage; n = 12 in S1, S2, S5, S6 (kEO = 29: the full form spills 100 bytes more at 128 registers and gains nothing,
// this one +6 %); n = 6 not at all.  The plain, the dot-fused and the xpay-fused kernel of an n use the same mask: same
// arithmetic, same bits.
template <int N> constexpr bool kUseEO = (N == 8 || N == 10 || N == 12);
constexpr int kEO12 = 29;

template <int N> int dispatch_ax_dot(int variant, bool eo, size_t E, const double *u, const double *g, double *w, cudaStream_t s, AxDotArgs dot) {
  constexpr int G = Shape<N>::G, W = Shape<N>::W, GPC = Shape<N>::GPC;
  constexpr int kThreads = GPC * W * 32;
  constexpr int MB168 = 65536 / (168 * kThreads) > 0 ? 65536 / (168 * kThreads) : 1;
  if constexpr (kUseEO<N>) {
    if (eo && variant == 0) {
      if constexpr (N == 8) return launch_ax<N, G, W, GPC, 2, 4, false, MB168, true, true, false, 0, 0, kEOAll>(E, u, g, w, s, dot);
      else if constexpr (N == 12) return launch_ax<N, G, W, GPC, 2, 6, false, (MB168 + 1), true, true, true, 4, 0, kEO12>(E, u, g, w, s, dot);
      // n = 10: with the even-odd stages the shape of the xpay-fused kernel (three buffers, three slabs in flight, 168
      // registers, two CTAs) beats the three-CTA shape for the fused dot product: 0.87 against 0.78 of the peak
      else return launch_ax<N, G, W, GPC, 3, 6, false, MB168, true, true, false, 3, 0, kEOAll>(E, u, g, w, s, dot);
    }
  }
  // kept for profiling (tools/ax_sweep.py axdot): round 1's three-buffer shape, two CTAs with the local window, three
  // CTAs with the wrapping window
  // the shapes of the xpay-fused kernel (three buffers, 168 registers, local window) for the dot product alone
  if (variant == 70 || variant == 71) {
    if constexpr (kUseEO<N>) {
      if (eo) {
        if (variant == 70) return launch_ax<N, G, W, GPC, 3, 6, false, MB168, true, true, false, 3, 0, kEOAll>(E, u, g, w, s, dot);
        return launch_ax<N, G, W, GPC, 2, 4, false, MB168, true, true, false, 3, 0, kEOAll>(E, u, g, w, s, dot);
      }
    }
    if (variant == 70) return launch_ax<N, G, W, GPC, 3, 6, false, MB168, true, true, false, 3>(E, u, g, w, s, dot);
    return launch_ax<N, G, W, GPC, 2, 4, false, MB168, true, true, false, 3>(E, u, g, w, s, dot);
  }
  if (variant == 63) return launch_ax<N, G, W, GPC, 2, 4, false, (MB168 + 1), true, true, true, 3, 1>(E, u, g, w, s, dot);
  if (variant == 64) return launch_ax<N, G, W, GPC, 2, 4, false, (MB168 + 1), true, true, true, 3, 0>(E, u, g, w, s, dot);
  if (variant == 67) return launch_ax<N, G, W, GPC, 2, 4, false, (MB168 + 1), true, true, true, 5, 1>(E, u, g, w, s, dot);
  if (variant == 60) return launch_ax<N, G, W, GPC, 2, 4, false, MB168, true>(E, u, g, w, s, dot);
  if (variant == 61) return launch_ax<N, G, W, GPC, 2, 4, false, MB168, true, true, true, 3>(E, u, g, w, s, dot);
  if (variant == 62) return launch_ax<N, G, W, GPC, 2, 4, false, (MB168 + 1), true, true, true, 0>(E, u, g, w, s, dot);
  // the shapes of dispatch_ax (variant 0), with the dot product
  if constexpr (N == 8) return launch_ax<N, G, W, GPC, 2, 4, false, MB168, true>(E, u, g, w, s, dot);
  else if constexpr (N == 12) return launch_ax<N, G, W, GPC, 2, 6, false, (MB168 + 1), true, true, true, 4>(E, u, g, w, s, dot);
  else if constexpr (N == 10) return launch_ax<N, G, W, GPC, 2, 4, false, (MB168 + 1), true, true, true, 3, 1>(E, u, g, w, s, dot);   // kPin
  else return launch_ax<N, G, W, GPC, 2, 4, false, (MB168 + 1), true, true, true, 3>(E, u, g, w, s, dot);
}

// p <- r + beta p fused in front (the shapes of dispatch_ax_dot).
template <int N> int dispatch_ax_xpay_dot(bool eo, size_t E, const double *g, double *w, cudaStream_t s, AxDotArgs dot, AxXpayArgs xp) {
  constexpr int G = Shape<N>::G, W = Shape<N>::W, GPC = Shape<N>::GPC;
  constexpr int kThreads = GPC * W * 32;
  constexpr int MB168 = 65536 / (168 * kThreads) > 0 ? 65536 / (168 * kThreads) : 1;
  if constexpr (kUseEO<N>) {
    if (eo) {
      if constexpr (N == 8) return launch_ax_xpay_dot<N, G, W, GPC, 2, 4, MB168, kEOAll>(E, g, w, s, dot, xp);
      else if constexpr (N == 12) return launch_ax_xpay_dot<N, G, W, GPC, 2, 4, MB168, kEO12>(E, g, w, s, dot, xp);
      else return launch_ax_xpay_dot<N, G, W, GPC, 3, 6, MB168, kEOAll>(E, g, w, s, dot, xp);
    }
  }
  if constexpr (N == 8 || N == 12) return launch_ax_xpay_dot<N, G, W, GPC, 2, 4, MB168>(E, g, w, s, dot, xp);   // n = 12: no spills this way
  else return launch_ax_xpay_dot<N, G, W, GPC, 3, 6, MB168>(E, g, w, s, dot, xp);
}

#endif

// The even-odd tables of this unit from D (device memory), on `stream`.
inline cudaError_t stage_eo(const double *D, int n, cudaStream_t stream) {
  static double *eo = nullptr;
  if (!eo) {
    cudaError_t err = cudaGetSymbolAddress(reinterpret_cast<void **>(&eo), nompk_ax_cEO);
    if (err != cudaSuccess) return err;
  }
  ax_stage_eo<<<1, 256, 0, stream>>>(D, eo, n);
  return cudaGetLastError();
}

#if NOMPK_AX_PART == 0
template <int N> int dispatch_ax(bool eo, size_t E, const double *u, const double *g, double *w, cudaStream_t s) {
  constexpr int G = Shape<N>::G, W = Shape<N>::W, GPC = Shape<N>::GPC;
  constexpr int kThreads = GPC * W * 32;
  constexpr int MB168 = 65536 / (168 * kThreads) > 0 ? 65536 / (168 * kThreads) : 1;
  if constexpr (kUseEO<N>) {
    if (eo) {
      if constexpr (N == 8) return launch_ax<N, G, W, GPC, 2, 4, false, MB168, false, true, false, 0, 0, kEOAll>(E, u, g, w, s);
      else if constexpr (N == 12) return launch_ax<N, G, W, GPC, 2, 6, false, (MB168 + 1), false, true, true, 4, 0, kEO12>(E, u, g, w, s);
      else return launch_ax<N, G, W, GPC, 2, 4, false, (MB168 + 1), false, true, true, 3, 0, kEOAll>(E, u, g, w, s);
    }
  }
  // production choice (interleaved sweeps of round 2, profiles/r02_kernel_sweeps.jsonl)
  // n = 8: one warp per element, three buffers, 168 registers, window running on into the next element (no over-read
  // at this element time).  The others: two shared buffers per element, one CTA more per SM (128 registers) and the
  // LOCAL prefetch window, which removed the 14 - 30 % of DRAM reads that the wrapping window fetched twice on these
  // shapes: n = 10 +6 %, n = 12 +9 %, n = 6 +13 % over the round-1 choices in the same interleaved sweeps.
  if constexpr (N == 8) return launch_ax<N, G, W, GPC, 2, 4, false, MB168>(E, u, g, w, s);
  else if constexpr (N == 12) return launch_ax<N, G, W, GPC, 2, 6, false, (MB168 + 1), false, true, true, 4>(E, u, g, w, s);
  else return launch_ax<N, G, W, GPC, 2, 4, false, (MB168 + 1), false, true, true, 3>(E, u, g, w, s);
}
#else
// Variants kept for profiling (tools/ax_sweep.py sweeps them).
template <int N> int dispatch_ax_variants(int variant, size_t E, const double *u, const double *g, double *w, cudaStream_t s) {
  // MBr = resident CTAs per SM that cap the kernel at r registers per thread.
  constexpr int G = Shape<N>::G, W = Shape<N>::W, GPC = Shape<N>::GPC;
  constexpr int kThreads = GPC * W * 32;
  constexpr int MB128 = 65536 / (128 * kThreads), MB168 = 65536 / (168 * kThreads) > 0 ? 65536 / (168 * kThreads) : 1;
  // one element per group of ceil(T / 32) warps, about 128 threads per CTA
  constexpr int W1 = Layout<N>::WPE, GPC1 = (128 / Layout<N>::LPE > 0 ? 128 / Layout<N>::LPE : 1), kThreads1 = GPC1 * W1 * 32;
  constexpr int MB128_1 = 65536 / (128 * kThreads1), MB168_1 = 65536 / (168 * kThreads1);
  switch (variant) {
#if NOMPK_AX_PART == 1
  case 1: return launch_ax<N, G, W, GPC, 2, 4, false, MB128>(E, u, g, w, s);
  case 2: return launch_ax<N, G, W, GPC, 2, 3, false, MB128>(E, u, g, w, s);
  case 3: return launch_ax<N, G, W, GPC, 2, 2, false, MB128>(E, u, g, w, s);
  case 4: return launch_ax<N, G, W, GPC, 2, 0, false, MB128>(E, u, g, w, s);
  case 5: return launch_ax<N, G, W, GPC, 2, 4, true, MB128>(E, u, g, w, s);
  case 6: return launch_ax<N, G, W, GPC, 2, 0, false, 1>(E, u, g, w, s);
  case 7: return launch_ax<N, G, W, GPC, 2, 4, false, MB168>(E, u, g, w, s);
  case 8: return launch_ax<N, G, W, GPC, 3, 6, false, MB168>(E, u, g, w, s);
  case 9: return launch_ax<N, G, W, GPC, 2, 4, false, 1>(E, u, g, w, s);
  case 10: return launch_ax<N, G, W, GPC, 4, 6, false, 1>(E, u, g, w, s);
  case 11: return launch_ax<N, G, W, GPC, 3, 4, false, MB168>(E, u, g, w, s);
  case 13: return launch_ax<N, G, W, GPC, 2, 4, false, MB168, false, false>(E, u, g, w, s);   // one CTA per group
  case 14: return launch_ax<N, G, W, GPC, 3, 6, false, MB168, false, false>(E, u, g, w, s);
  case 15: return launch_ax<N, G, W, GPC, 2, 4, false, MB128, false, false>(E, u, g, w, s);
  case 16: return launch_ax<N, G, W, GPC, 2, 0, false, MB168, false, false>(E, u, g, w, s);
  case 17: return launch_ax<N, G, W, GPC, 2, 8, false, MB168, false, false>(E, u, g, w, s);
  // experimental two-buffer variants (see ax_kernel): 168 registers / 2 CTAs, and 3 CTAs per SM where the registers fit
  case 21: return launch_ax<N, G, W, GPC, 2, 4, false, MB168, false, true, true>(E, u, g, w, s);
  case 22: return launch_ax<N, G, W, GPC, 2, 4, false, (MB168 + 1), false, true, true>(E, u, g, w, s);
  case 23: return launch_ax<N, G, W, GPC, 1, 4, false, (MB168 + 1), false, true, true>(E, u, g, w, s);  // one slab in flight
  case 12:  // the one-element-on-ceil(T/32)-warps shape of the first version, for comparison
    return launch_ax<N, 1, Layout<N>::WPE, (128 / Layout<N>::LPE > 0 ? 128 / Layout<N>::LPE : 1), 2, 4, false, 1>(E, u, g, w, s);
#else
  // round 2: shape and eviction priority of the prefetch window (kPfMode) on the three-CTA / two-buffer shapes ...
  case 30: return launch_ax<N, G, W, GPC, 2, 4, false, (MB168 + 1), false, true, true, 3>(E, u, g, w, s);
  case 31: return launch_ax<N, G, W, GPC, 2, 4, false, (MB168 + 1), false, true, true, 4>(E, u, g, w, s);
  case 32: return launch_ax<N, G, W, GPC, 2, 6, false, (MB168 + 1), false, true, true, 3>(E, u, g, w, s);
  case 33: return launch_ax<N, G, W, GPC, 2, 6, false, (MB168 + 1), false, true, true, 4>(E, u, g, w, s);
  case 34: return launch_ax<N, G, W, GPC, 1, 4, false, (MB168 + 1), false, true, true, 3>(E, u, g, w, s);
  case 35: return launch_ax<N, G, W, GPC, 1, 4, false, (MB168 + 1), false, true, true, 4>(E, u, g, w, s);
  case 36: return launch_ax<N, G, W, GPC, 1, 6, false, (MB168 + 1), false, true, true, 3>(E, u, g, w, s);
  case 37: return launch_ax<N, G, W, GPC, 2, 8, false, (MB168 + 1), false, true, true, 3>(E, u, g, w, s);
  case 38: return launch_ax<N, G, W, GPC, 2, 4, false, (MB168 + 1), false, true, true, 2>(E, u, g, w, s);
  // ... and on the three-buffer / two-CTA shapes
  case 39: return launch_ax<N, G, W, GPC, 2, 4, false, MB168, false, true, false, 3>(E, u, g, w, s);
  case 40: return launch_ax<N, G, W, GPC, 2, 6, false, MB168, false, true, false, 3>(E, u, g, w, s);
  case 41: return launch_ax<N, G, W, GPC, 3, 6, false, MB168, false, true, false, 3>(E, u, g, w, s);
  // ... and one element per ceil(T / 32) warps (n = 10: 2 warps, 50 of 64 lanes; n = 12: 3 warps, 72 of 96): fewer lanes do
  // useful work, but the groups are small and many -- barriers of 2 - 3 warps, 6 - 8 independent groups per SM
  case 42: return launch_ax<N, 1, W1, GPC1, 2, 4, false, MB168_1, false, true, false, 3>(E, u, g, w, s);
  case 43: return launch_ax<N, 1, W1, GPC1, 2, 4, false, MB128_1, false, true, true, 3>(E, u, g, w, s);
  case 44: return launch_ax<N, 1, W1, GPC1, 1, 4, false, MB128_1, false, true, true, 3>(E, u, g, w, s);
  case 45: return launch_ax<N, 1, W1, GPC1, 2, 4, false, MB128_1, false, true, true, 4>(E, u, g, w, s);
  case 46: return launch_ax<N, 1, W1, GPC1, 3, 6, false, MB168_1, false, true, false, 3>(E, u, g, w, s);
  case 47: return launch_ax<N, 1, W1, GPC1, 2, 4, false, MB168_1, false, true, true, 3>(E, u, g, w, s);
  // ... the first slabs of an element prefetched too (kPfMode 5), and the loads that fill the register ring pinned in
  // front of the barrier of S4 (kPin)
  case 48: return launch_ax<N, G, W, GPC, 2, 4, false, (MB168 + 1), false, true, true, 5>(E, u, g, w, s);
  case 50: return launch_ax<N, G, W, GPC, 2, 6, false, (MB168 + 1), false, true, true, 5>(E, u, g, w, s);
  case 52: return launch_ax<N, G, W, GPC, 2, 4, false, (MB168 + 1), false, true, true, 3, 1>(E, u, g, w, s);
  // even-odd contractions (kEO) on the production shapes; D must be centro-antisymmetric
  case 54:
    if constexpr (N == 8) return launch_ax<N, G, W, GPC, 2, 4, false, MB168, false, true, false, 0, 0, kEOAll>(E, u, g, w, s);
    else if constexpr (N == 12) return launch_ax<N, G, W, GPC, 2, 6, false, (MB168 + 1), false, true, true, 4, 0, kEOAll>(E, u, g, w, s);
    else return launch_ax<N, G, W, GPC, 2, 4, false, (MB168 + 1), false, true, true, 3, 0, kEOAll>(E, u, g, w, s);
  case 55:   // ... with the 168-register budget (two CTAs per SM for n = 10, 12)
    return launch_ax<N, G, W, GPC, 2, 4, false, MB168, false, true, true, 3, 0, kEOAll>(E, u, g, w, s);
  case 56: return launch_ax<N, G, W, GPC, 2, 4, false, MB168, false, true, false, 3, 0, kEOAll>(E, u, g, w, s);   // ... three buffers
  case 57: return launch_ax<N, G, W, GPC, 3, 6, false, MB168, false, true, false, 3, 0, kEOAll>(E, u, g, w, s);   // ... three slabs in flight
  // ... in some of the stages only (n = 12: the full form spills): S1, S5, S6 / S1, S2, S5, S6, on n = 12's production shape
  case 58: return launch_ax<N, G, W, GPC, 2, 6, false, (MB168 + 1), false, true, true, 4, 0, 25>(E, u, g, w, s);
  case 59: return launch_ax<N, G, W, GPC, 2, 6, false, (MB168 + 1), false, true, true, 4, 0, 29>(E, u, g, w, s);
#endif
  default: return NOMPK_AX_NO_SUCH_VARIANT;
  }
}
#endif

}  // namespace
}  // namespace nompk

#define NOMPK_AX_VARIANTS_DECL_(part, n) NOMPK_AX_VARIANTS_DECL(part, n)
#if NOMPK_AX_PART == 0
NOMPK_AX_VARIANTS_DECL_(1, NOMPK_AX_N);
NOMPK_AX_VARIANTS_DECL_(2, NOMPK_AX_N);
#define NOMPK_AX_CALL_VARIANTS__(part, n) nompk_ax_variants##part##_n##n
#define NOMPK_AX_CALL_VARIANTS_(part, n) NOMPK_AX_CALL_VARIANTS__(part, n)
#define NOMPK_AX_RUN_DEFINE_(n) NOMPK_AX_RUN_DECL(n)
#define NOMPK_AX_RUN_DEFINE(n) NOMPK_AX_RUN_DEFINE_(n)
NOMPK_AX_RUN_DEFINE(NOMPK_AX_N) {
  using namespace nompk;
  constexpr int n = NOMPK_AX_N;
  if (!dot && variant != 0) {   // a profiling shape?  (unknown numbers run the production kernel)
    int rc = NOMPK_AX_CALL_VARIANTS_(1, NOMPK_AX_N)(variant, E, u, g, D, w, stream);
    if (rc == NOMPK_AX_NO_SUCH_VARIANT) rc = NOMPK_AX_CALL_VARIANTS_(2, NOMPK_AX_N)(variant, E, u, g, D, w, stream);
    if (rc != NOMPK_AX_NO_SUCH_VARIANT) return rc;
  }
  const bool eo = kUseEO<n> && (flags & NOMPK_AX_D_ANTISYMMETRIC);
  if (!(flags & NOMPK_AX_D_CACHED)) {
    NOMPK_CUDA_TRY(cudaMemcpyToSymbolAsync(nompk_ax_cD, D, sizeof(double) * n * n, 0, cudaMemcpyDeviceToDevice, stream));
    if (eo) NOMPK_CUDA_TRY(stage_eo(D, n, stream));
  }
  if (dot && xpay)
    return dispatch_ax_xpay_dot<n>(eo, E, g, w, stream, *static_cast<const AxDotArgs *>(dot), *static_cast<const AxXpayArgs *>(xpay));
  if (dot) return dispatch_ax_dot<n>(variant, eo, E, u, g, w, stream, *static_cast<const AxDotArgs *>(dot));
  return dispatch_ax<n>(eo, E, u, g, w, stream);
}
#else
NOMPK_AX_VARIANTS_DECL_(NOMPK_AX_PART, NOMPK_AX_N) {
  using namespace nompk;
  constexpr int n = NOMPK_AX_N;
  if (!((NOMPK_AX_PART == 1 && ((variant >= 1 && variant <= 17) || (variant >= 21 && variant <= 23))) ||
        (NOMPK_AX_PART == 2 && variant >= 30 && variant <= 59)))
    return NOMPK_AX_NO_SUCH_VARIANT;   // before D is staged for nothing
  NOMPK_CUDA_TRY(cudaMemcpyToSymbolAsync(nompk_ax_cD, D, sizeof(double) * n * n, 0, cudaMemcpyDeviceToDevice, stream));
  if (variant >= 54) NOMPK_CUDA_TRY(stage_eo(D, n, stream));
  return dispatch_ax_variants<n>(variant, E, u, g, w, stream);
}
#endif

#else  // NOMPK_AX_N == 0: the C ABI

namespace nompk {
namespace {
int g_variant = 0;
}
}  // namespace nompk

extern "C" int nompk_ax_supported(int n) { return n == 8 || n == 10 || n == 6 || n == 12; }

extern "C" int nompk_ax_set_variant(int variant) {
  nompk::g_variant = variant;
  return NOMPK_OK;
}

static int ax_common(int n, size_t E, const double *u, const double *g, const double *D, double *w, unsigned flags,
                     cudaStream_t stream, const nompk::AxDotArgs *dot, const nompk::AxXpayArgs *xpay = nullptr);

extern "C" int nompk_ax_f64(int n, size_t E, const double *u, const double *g, const double *D, double *w,
                            unsigned flags, void *stream_) {
  return ax_common(n, E, u, g, D, w, flags, static_cast<cudaStream_t>(stream_), nullptr);
}

extern "C" int nompk_ax_dot_f64(int n, size_t E, const double *u, const double *g, const double *D, double *w,
                                double *result, double *result_host_mapped, unsigned long long host_seq,
                                void *workspace, unsigned flags, void *stream_) {
  return nompk_ax_dot_peers_f64(n, E, u, g, D, w, result, result_host_mapped, host_seq, workspace, nullptr, flags, stream_);
}

extern "C" int nompk_ax_dot_peers_f64(int n, size_t E, const double *u, const double *g, const double *D, double *w,
                                      double *result, double *result_host_mapped, unsigned long long host_seq,
                                      void *workspace, const nompk_peers_t *peers, unsigned flags, void *stream_) {
  using namespace nompk;
  if (!result || !workspace) {
    set_error("nompk_ax_dot_f64: NULL result/workspace");
    return NOMPK_EINVAL;
  }
  AxDotArgs dot;
  dot.workspace = workspace;
  dot.result = result, dot.result_host = result_host_mapped, dot.host_seq = host_seq;
  if (peers && peers->world > 1) {
    if (!dot.px.set(peers, kMaxFusedRanks)) {
      set_error("nompk_ax_dot_peers_f64: bad peer description (rank %d of %d; at most %d ranks)", peers->rank, peers->world,
                kMaxFusedRanks);
      return NOMPK_EINVAL;
    }
    if (E == 0) {
      set_error("nompk_ax_dot_peers_f64: every rank needs at least one element");
      return NOMPK_EINVAL;
    }
  }
  if (E == 0) {  // identity, through the same publication protocol
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NOMPK_CUDA_TRY(cudaMemsetAsync(result, 0, sizeof(double), stream));
    if (result_host_mapped) {
      NOMPK_CUDA_TRY(cudaMemsetAsync(result_host_mapped, 0, sizeof(double), stream));
      NOMPK_CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<char *>(result_host_mapped) + 8, &dot.host_seq, 8,
                                     cudaMemcpyHostToDevice, stream));
      NOMPK_CUDA_TRY(cudaStreamSynchronize(stream));
    }
    return NOMPK_OK;
  }
  return ax_common(n, E, u, g, D, w, flags, static_cast<cudaStream_t>(stream_), &dot);
}

extern "C" int nompk_ax_xpay_dot_peers_f64(int n, size_t E, double *p, const double *r, double beta, const double *beta_dev,
                                           const double *g, const double *D, double *w, double *result,
                                           double *result_host_mapped, unsigned long long host_seq, void *workspace,
                                           const nompk_peers_t *peers, unsigned flags, void *stream_) {
  using namespace nompk;
  if (!result || !workspace || !r || !is_aligned16(r)) {
    set_error("nompk_ax_xpay_dot_peers_f64: NULL result / workspace / r, or r not 16-byte aligned");
    return NOMPK_EINVAL;
  }
  AxDotArgs dot;
  dot.workspace = workspace;
  dot.result = result, dot.result_host = result_host_mapped, dot.host_seq = host_seq;
  if (peers && peers->world > 1) {
    if (E == 0 || !dot.px.set(peers, kMaxFusedRanks)) {
      set_error("nompk_ax_xpay_dot_peers_f64: bad peer description (rank %d of %d), or a rank without elements", peers->rank,
                peers->world);
      return NOMPK_EINVAL;
    }
  }
  if (E == 0)  // nothing to update; the dot product is the identity, through the same publication protocol
    return nompk_ax_dot_peers_f64(n, 0, p, g, D, w, result, result_host_mapped, host_seq, workspace, nullptr, flags, stream_);
  AxXpayArgs xp;
  xp.r = r, xp.p = p, xp.beta_dev = beta_dev, xp.beta = beta;
  return ax_common(n, E, p, g, D, w, flags, static_cast<cudaStream_t>(stream_), &dot, &xp);
}

static int ax_common(int n, size_t E, const double *u, const double *g, const double *D, double *w, unsigned flags,
                     cudaStream_t stream, const nompk::AxDotArgs *dot, const nompk::AxXpayArgs *xpay) {
  using namespace nompk;
  if (!nompk_ax_supported(n)) {
    set_error("nompk_ax_f64: n = %d has no hand-written kernel (supported: 6, 8, 10, 12)", n);
    return NOMPK_EUNSUPPORTED;
  }
  if (E == 0) return NOMPK_OK;
  if (!u || !g || !D || !w) {
    set_error("nompk_ax_f64: NULL operand");
    return NOMPK_EINVAL;
  }
  if (!is_aligned16(u) || !is_aligned16(g) || !is_aligned16(w)) {
    set_error("nompk_ax_f64: u, g and w must be 16-byte aligned");
    return NOMPK_EINVAL;
  }
  switch (n) {
  case 6: return nompk_ax_run_n6(g_variant, E, u, g, D, w, flags, stream, dot, xpay);
  case 8: return nompk_ax_run_n8(g_variant, E, u, g, D, w, flags, stream, dot, xpay);
  case 10: return nompk_ax_run_n10(g_variant, E, u, g, D, w, flags, stream, dot, xpay);
  case 12: return nompk_ax_run_n12(g_variant, E, u, g, D, w, flags, stream, dot, xpay);
  }
  return NOMPK_EUNSUPPORTED;
}

#endif  // NOMPK_AX_N
