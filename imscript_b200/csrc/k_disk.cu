// k_disk.cu -- erosion / dilation by convex symmetric "row-run" elements
// (disks): BASELINE configs C2 (disk7 opening/closing) and C4 (disk15 tophat).
//
// A disk of reach R is one centred run per row, half-width hw(dy).  The kernel
// is a streaming register march over a column strip and a row band:
//   * the strip's input rows (plus halo columns) stream through a shared-memory
//     ring of row GROUPS.  One elected lane issues one TMA tensor copy
//     (cp.async.bulk.tensor.4d -> UTMALDG) per group, a group ahead of the
//     compute; completion is signalled on an mbarrier per ring slot, the
//     consumer warps hand the slot back through a second mbarrier -- one barrier
//     round trip per group, not per row.  The tensor map views a plane as
//     (bw, w/bw, rows, planes) so that a box row of any width is one dense
//     shared-memory row, and fills everything outside the image with NaN
//     (CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA), which min.f32 /
//     max.f32 ignore -- exactly the reference's getpixel_nan border rule
//     (src/morsi.c:30-35), done by the copy engine;
//   * a thread owns C adjacent columns and walks down the band two rows at a
//     time.  For each new row it builds the nested horizontal running extrema
//     H_k(x) = ext(in[x-k..x+k]) with one 3-input FMNMX3 per k, and folds
//     H_{hw(dy)} of the row pair into the 2R+2 register accumulators of the
//     output rows the pair touches (one FMNMX3 per accumulator and pair).  The
//     2M+1 centre rows of a disk share one half-width, so their contribution is
//     folded once from pair / quad / octet extrema that all outputs share.
//     Cost per sample and stage: about RX + (R - M) + 2.5 FMNMX instead of the
//     n = e[0] fmin/fmax calls of src/morsi.c:65 (12 vs 145 for disk7, 26.5 vs
//     697 for disk15);
//   * two-stage operations (opening, closing, tophat, bothat) run both stages
//     in one CTA: the first half of the warps reduces the input ring into a
//     small shared-memory ring of temporary rows, the second half reduces that
//     ring and writes the result -- the temporary image of src/morsi.c:143-146
//     never exists in HBM (4 B read + 4 B written per sample).  The first
//     stage stores its rows NEGATED: max(t) = -min(-t), so both halves run the
//     same instruction stream (one copy in the instruction cache).  Temporary
//     samples outside the image are NaN (absent), SURVEY.md 9.1-B.
// The accumulator ring is addressed at compile time by unrolling R+1 steps, so
// the element's shape is a template parameter (shapes.cuh).
//
// min.f32/max.f32 do not keep the reference's last-wins order for +0/-0:
// the kernel raises *flag when it sees a -0.0 and the dispatcher re-runs the
// order-preserving path.
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include <cuda.h>

#include "dispatch.cuh"
#include "shapes.cuh"

struct DiskArgs {
	Band src;          // image the (first) reduction runs over
	Band xop;          // x operand of the epilogue (may be unused)
	Band other;        // second reduction operand of the epilogue (single stage only)
	float *y;
	long long y_pstride;
	int y_row0, y_rows;
	int src_rows;      // rows held by src (for the loader's bounds)
	int w, h;
	int pitch;         // floats between rows of every operand (>= w, multiple of 4)
	int band_rows;     // output rows per CTA
	int epi;
	int *flag;
	int both = 0;      // 1: erosion and dilation in one pass (k_disk_both), epilogue epi(a, b, x)
};

// ---- small PTX wrappers ------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(unsigned bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
// try_wait suspends the thread for a hardware time slice; the loop is left when
// the phase completes.  A wait that lasts seconds is a protocol bug (a legal one
// lasts microseconds): trap, so that it surfaces as a launch failure instead of
// a hung device.
__device__ __forceinline__ unsigned mbar_try(unsigned bar, unsigned parity)
{
	unsigned ok;
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
		"selp.u32 %0, 1, 0, p;\n"
		"}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
	return ok;
}
static __device__ __noinline__ void mbar_wait_slow(unsigned bar, unsigned parity)
{
	const long long t0 = clock64();
	while (!mbar_try(bar, parity))
		if (clock64() - t0 > 8000000000LL) __trap();
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
	if (!mbar_try(bar, parity)) mbar_wait_slow(bar, parity);
}
// Wait executed by a whole (converged) warp.  The lanes of a warp can leave the
// try_wait loop in different iterations; nothing would reconverge them
// afterwards, and the lanes that fell behind would keep reading ring slots
// that lane 0 -- which arrives for the warp -- has already handed back.
__device__ __forceinline__ void mbar_wait_warp(unsigned bar, unsigned parity)
{
	mbar_wait(bar, parity);
	__syncwarp();
}
// one group of rows of the strip, global -> shared, by the TMA engine; completion on `bar`
__device__ __forceinline__ void tma_load_4d(unsigned dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, unsigned bar)
{
	asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
		:: "r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

template <bool ISMAX> __device__ __forceinline__ float dext2(float a, float b)
{
	return ISMAX ? fmaxf(a, b) : fminf(a, b);
}
template <bool ISMAX> __device__ __forceinline__ float dext3(float a, float b, float c)
{
	return ISMAX ? fmaxf(fmaxf(a, b), c) : fminf(fminf(a, b), c);   // one FMNMX3
}

__device__ __forceinline__ float disk_epi(int epi, float a, float b, float x)
{
	switch (epi) {
	case EPI_A: return a;
	case EPI_B: return b;
	case EPI_B_SUB_A: return epilogue<EPI_B_SUB_A>(a, b, x);
	case EPI_X_SUB_A: return epilogue<EPI_X_SUB_A>(a, b, x);
	case EPI_B_SUB_X: return epilogue<EPI_B_SUB_X>(a, b, x);
	case EPI_LAP: return epilogue<EPI_LAP>(a, b, x);
	case EPI_ENH: return epilogue<EPI_ENH>(a, b, x);
	case EPI_BLUR: return epilogue<EPI_BLUR>(a, b, x);
	case EPI_A_SUB_B: return epilogue<EPI_A_SUB_B>(a, b, x);
	case EPI_X_SUB_B: return epilogue<EPI_X_SUB_B>(a, b, x);
	case EPI_A_SUB_X: return epilogue<EPI_A_SUB_X>(a, b, x);
	case EPI_IBLUR: return epilogue<EPI_IBLUR>(a, b, x);
	case EPI_EBLUR: return epilogue<EPI_EBLUR>(a, b, x);
	case EPI_CBLUR: return epilogue<EPI_CBLUR>(a, b, x);
	}
	return a;
}

// Epilogue with extra operands (single-stage kernel), kept out of line: the
// unrolled march calls it from 2(R+1) places.  Arithmetic only: the operands
// are loaded by the caller BEFORE the step's reductions, so that their latency
// hides behind them (a load inside this call would stall every step).
template <bool ISMAX>
__device__ __noinline__ float4 disk_epi4(int epi, float m0, float m1, float m2, float m3,
		float o0, float o1, float o2, float o3, float x0, float x1, float x2, float x3)
{
	const float m[4] = {m0, m1, m2, m3}, ov[4] = {o0, o1, o2, o3}, xv[4] = {x0, x1, x2, x3};
	float out[4] = {m0, m1, m2, m3};
	// one dispatch per call, the four columns inside each case
#define CASE(E) case E: _Pragma("unroll") for (int c = 0; c < 4; c++) \
		out[c] = ISMAX ? epilogue<E>(ov[c], m[c], xv[c]) : epilogue<E>(m[c], ov[c], xv[c]); break;
	switch (epi) {
	CASE(EPI_B_SUB_A) CASE(EPI_X_SUB_A) CASE(EPI_B_SUB_X) CASE(EPI_LAP) CASE(EPI_ENH) CASE(EPI_BLUR)
	CASE(EPI_A_SUB_B) CASE(EPI_X_SUB_B) CASE(EPI_A_SUB_X) CASE(EPI_IBLUR) CASE(EPI_EBLUR) CASE(EPI_CBLUR)
	default: break;                                  // EPI_A / EPI_B: the reduction itself
	}
#undef CASE
	return make_float4(out[0], out[1], out[2], out[3]);
}

// ---- compile-time geometry ---------------------------------------------------------
template <class S> struct DiskShape {
	static constexpr int R = S::R;
	static constexpr int RX = S::hw(R);               // half-width of the centre row
	// the centre run: rows R-M .. R+M all have half-width RX
	__host__ __device__ static constexpr int centre()
	{
		int m = 0;
		while (m < R && S::hw(R - m - 1) == RX && S::hw(R + m + 1) == RX) m++;
		return m;
	}
	static constexpr int M = centre();
	// pair/quad/octet sharing pays from 3 full centre pairs on
	static constexpr bool SHARE = M >= 3 && M <= 5 && M < R;
};

template <class S, int C, int W, bool TWO>
struct DiskCfg {
	static constexpr int R = S::R;
	static constexpr int RX = DiskShape<S>::RX;
	// A thread's window starts LH columns left of its own columns; two stages
	// need an even LH (their windows nest with 16-byte alignment).
	static constexpr int LH = TWO ? (RX + 1) / 2 * 2 : (RX + 3) / 4 * 4;
	static constexpr int NV = (C + LH + RX + C - 1) / C * C;    // floats a thread reads per row
	static constexpr int NT = 32 * W;                 // threads per stage
	static constexpr int TW = NT * C;                 // columns a stage produces per row
	static constexpr int OUTW = TWO ? (TW - LH - RX) / 4 * 4 : TW;   // output columns per strip
	// ring row: the windows of all threads (TW + NV - C floats) plus the slack
	// that lets the TMA box start on a bw <= 32 column boundary; whole 128 bytes
	static constexpr int RP = (TW + NV - C + 28 + 31) / 32 * 32;
	// The march is unrolled over PERIOD steps (a multiple of R+1, the period of
	// the accumulator ring).  Rows travel in groups of GP pairs (GP divides
	// PERIOD, so a step's place in its group is a compile-time constant).
	static constexpr int PERIOD = (R + 1) * ((6 + R) / (R + 1));
	__host__ __device__ static constexpr int group_pairs()
	{
		int best = 1;
		for (int d = 1; d <= PERIOD; d++)
			if (PERIOD % d == 0 && d * 2 * RP * 4 <= 18 * 1024) best = d;
		return best;
	}
	static constexpr int GP = group_pairs();          // row pairs per group
	static constexpr int PAIR = 2 * RP;               // floats per row pair
	static constexpr int GROUP = GP * PAIR;           // floats per group
	// groups per ring: two, or more when the groups are small (a prime period
	// leaves groups of one pair: keep ~20 KB in flight all the same)
#ifndef DISK_RING_BYTES
#define DISK_RING_BYTES (20 * 1024)
#endif
	// the 255-register shapes run one CTA per SM: shared memory is plentiful there, and a deeper
	// ring gives the two stages (which otherwise meet at every group boundary) room to drift
#ifndef DISK_RING_BYTES_BIG
#define DISK_RING_BYTES_BIG (20 * 1024)
#endif
	// measured on B200 (profiles/r2_kdisk_ring_variants.txt): three groups help the single-stage kernel
	// (disk7 erosion of 4096x4096x3: 0.085 -> 0.079 ms = 5.1 TB/s, still three CTAs per SM); the fused
	// two-stage kernel loses with them (one CTA per SM less: 0.149 -> 0.174 ms), and so do the big disks
#ifndef DISK_RING_BYTES_SINGLE
#define DISK_RING_BYTES_SINGLE (54 * 1024)
#endif
	static constexpr int RING_BYTES = (C * (2 * R + 2) > 64) ? DISK_RING_BYTES_BIG : (TWO ? DISK_RING_BYTES : DISK_RING_BYTES_SINGLE);
	static constexpr int NG = RING_BYTES / (GROUP * 4) < 2 ? 2 : (RING_BYTES / (GROUP * 4) > 8 ? 8 : RING_BYTES / (GROUP * 4));
	static constexpr unsigned GROUP_BYTES = GROUP * 4u;
	static constexpr int NACC = 2 * R + 2;            // accumulator slots per column
	static constexpr int THREADS = TWO ? 2 * NT : NT;
	static constexpr size_t SMEM = (size_t)NG * GROUP * (TWO ? 2 : 1) * sizeof(float)
		+ 4 * NG * sizeof(unsigned long long) + 128;
	// registers: the small disks need ~160 (3 CTAs of 128 threads per SM), the big ones all 255
#ifndef DISK_REGS_SMALL
#define DISK_REGS_SMALL 168
#endif
#ifndef DISK_REGS_C2
#define DISK_REGS_C2 84
#endif
#ifndef DISK_REGS_SMALL_W4
#define DISK_REGS_SMALL_W4 128
#endif
	static constexpr int REGS = (C * (2 * R + 2) > 64) ? 255 : (C == 2 ? (C * (2 * R + 2) > 36 ? 128 : DISK_REGS_C2) :
		(THREADS >= 256 ? DISK_REGS_SMALL_W4 : DISK_REGS_SMALL));
	static constexpr int MINB = 65536 / (REGS * THREADS) > 0 ? 65536 / (REGS * THREADS) : 1;
};

// ---- the per-thread march ------------------------------------------------------------
template <class S, int C, int LH, bool ISMAX>
struct DiskMarch {
	using G = DiskShape<S>;
	static constexpr int R = G::R, RX = G::RX, M = G::M;
	static constexpr bool SHARE = G::SHARE;
	static constexpr int NV = (C + LH + RX + C - 1) / C * C;
	static constexpr int NACC = 2 * R + 2;

	__device__ static __forceinline__ float init() { return ISMAX ? -CUDART_INF_F : CUDART_INF_F; }

	// nested horizontal extrema of one ring row; `base` points at the window's first float
	__device__ static __forceinline__ void chain(const float *base, float (&H)[RX + 1][C], int &zmin, bool check0)
	{
		float v[NV];
#pragma unroll
		for (int q = 0; q < NV / C; q++) {
			if (C == 4) { float4 t = *(const float4 *)(base + 4 * q); v[4*q] = t.x; v[4*q+1] = t.y; v[4*q+2] = t.z; v[4*q+3] = t.w; }
			else { float2 t = *(const float2 *)(base + 2 * q); v[2*q] = t.x; v[2*q+1] = t.y; }
		}
		// -0.0 is INT_MIN as a signed word: one 3-input integer min per two samples
		if (check0) {
#pragma unroll
			for (int c = 0; c < C; c += 2)
				zmin = min(min(zmin, __float_as_int(v[LH + c])), __float_as_int(v[LH + c + 1]));
		}
#pragma unroll
		for (int c = 0; c < C; c++) {
			H[0][c] = v[LH + c];
#pragma unroll
			for (int k = 1; k <= RX; k++)
				H[k][c] = dext3<ISMAX>(H[k - 1][c], v[LH + c - k], v[LH + c + k]);
		}
	}

	// One step = march rows i0 = 2g and i1 = 2g+1 (s = g mod (R+1); the caller's
	// loop over s is fully unrolled, so every index below folds to a constant).
	// They touch outputs o = i0 + d, d in [-2R, 1]; row i0 is row j0 = -d of
	// output o's window, row i1 is row j1 = 1-d.  The accumulator slot of o is
	// (2s + d) mod NACC.  Outputs i0-2R and i0-2R+1 are complete afterwards.
	// hs[c][]: P1 (pair extremum of the previous step), Q1, Q2 (quads ending
	// one / two steps ago), O1 (octet ending one step ago).
	template <class Release, class Emit>
	__device__ static __forceinline__ void step(float (&acc)[C][NACC], float (&hs)[C][4], const int s,
			const float *rowA, const float *rowB, int &zmin, bool check0, Release &&release, Emit &&emit)
	{
		float H0[RX + 1][C], H1[RX + 1][C];
		chain(rowA, H0, zmin, check0);
		chain(rowB, H1, zmin, check0);
		release();                     // the ring slot has been read
		float P0[C], Q0[C], O0[C];
		if (SHARE) {
#pragma unroll
			for (int c = 0; c < C; c++) {
				// the start value rides along: an all-NaN window must give +-INF
				// (src/morsi.c:63,77), and every output holds a centre pair
				P0[c] = dext3<ISMAX>(H0[RX][c], H1[RX][c], init());
				Q0[c] = dext2<ISMAX>(hs[c][0], P0[c]);
				O0[c] = M == 5 ? dext2<ISMAX>(hs[c][2], Q0[c]) : 0.f;
			}
		}
#pragma unroll
		for (int d = -2 * R; d <= 1; d++) {
			const int slot = ((2 * s + d) % NACC + NACC) % NACC;
			const int j0 = -d, j1 = 1 - d;
			const bool in0 = j0 >= 0, in1 = j1 <= 2 * R;
			const int k0 = S::hw(j0 < 0 ? 0 : j0);
			const int k1 = S::hw(j1 > 2 * R ? 2 * R : j1);
			const bool centre_pair = SHARE && in0 && in1 && j0 >= R - M && j1 <= R + M;
			const bool last_centre_pair = centre_pair && j1 + 2 > R + M;
#pragma unroll
			for (int c = 0; c < C; c++) {
				if (centre_pair) {
					if (last_centre_pair) {
						if (M == 3) acc[c][slot] = dext3<ISMAX>(acc[c][slot], hs[c][1], P0[c]);
						if (M == 4) acc[c][slot] = dext3<ISMAX>(acc[c][slot], hs[c][2], Q0[c]);
						if (M == 5) acc[c][slot] = dext3<ISMAX>(acc[c][slot], hs[c][3], P0[c]);
					}
				} else if (!in0) {
					acc[c][slot] = SHARE ? H1[k1][c] : dext2<ISMAX>(H1[k1][c], init());
				} else if (j0 == 0) {
					acc[c][slot] = dext3<ISMAX>(H0[k0][c], H1[k1][c], init());
				} else if (!in1) {
					acc[c][slot] = dext2<ISMAX>(acc[c][slot], H0[k0][c]);
				} else {
					acc[c][slot] = dext3<ISMAX>(acc[c][slot], H0[k0][c], H1[k1][c]);
				}
			}
		}
		if (SHARE) {
#pragma unroll
			for (int c = 0; c < C; c++) {
				hs[c][3] = O0[c];
				hs[c][2] = hs[c][1];
				hs[c][1] = Q0[c];
				hs[c][0] = P0[c];
			}
		}
		float m0[C], m1[C];
#pragma unroll
		for (int c = 0; c < C; c++) {
			m0[c] = acc[c][((2 * s - 2 * R) % NACC + NACC) % NACC];
			m1[c] = acc[c][((2 * s - 2 * R + 1) % NACC + NACC) % NACC];
		}
		emit(m0, m1);
	}
};

template <int C>
__device__ __forceinline__ void store_cols(float *q, const float (&m)[C])
{
	if (C == 4) *(float4 *)q = make_float4(m[0], m[1], m[2], m[3]);
	else *(float2 *)q = make_float2(m[0], m[1]);
}
template <int C>
__device__ __forceinline__ void load_cols(const float *q, float (&m)[C])
{
	if (C == 4) { float4 t = __ldg((const float4 *)q); m[0] = t.x; m[1] = t.y; m[2] = t.z; m[3] = t.w; }
	else { float2 t = __ldg((const float2 *)q); m[0] = t.x; m[1] = t.y; }
}

// first stage, rare path: a temporary row pair that touches the image border
// (scalars by value: arrays by reference would force the caller's rows into
// local memory on the common path too)
static __device__ __noinline__ void disk_store_tmp_masked(float *q0, float *q1, int ncol, unsigned colmask, bool row0_ok, bool row1_ok,
		float a0, float a1, float a2, float a3, float b0, float b1, float b2, float b3)
{
	const float nan = CUDART_NAN_F;
	a0 = (colmask & 1u) && row0_ok ? a0 : nan; b0 = (colmask & 1u) && row1_ok ? b0 : nan;
	a1 = (colmask & 2u) && row0_ok ? a1 : nan; b1 = (colmask & 2u) && row1_ok ? b1 : nan;
	a2 = (colmask & 4u) && row0_ok ? a2 : nan; b2 = (colmask & 4u) && row1_ok ? b2 : nan;
	a3 = (colmask & 8u) && row0_ok ? a3 : nan; b3 = (colmask & 8u) && row1_ok ? b3 : nan;
	if (ncol == 4) { *(float4 *)q0 = make_float4(a0, a1, a2, a3); *(float4 *)q1 = make_float4(b0, b1, b2, b3); }
	else { *(float2 *)q0 = make_float2(a0, a1); *(float2 *)q1 = make_float2(b0, b1); }
}

// ---- the kernel ----------------------------------------------------------------------
// TWO: warps [0,W) run the first reduction over the input ring, warps [W,2W)
// the second one over the temporary ring (ISMAX is the FIRST stage's flavour,
// and -- through the negated temporaries -- the flavour both halves compute).
// HASX (TWO only): the epilogue subtracts: tophat x - opening (ISMAX = false),
// bothat closing - x (ISMAX = true); src/morsi.c:229-245.
// tm: the source band as a (bw, w/bw, rows, planes) tensor, box (bw, RP/bw, 2*GP, 1).
// Which half of the warps runs which stage alternates between the CTAs an SM
// receives (a per-SM counter): a warp's scheduler partition is its index modulo
// 4, so with a fixed assignment every partition would only ever host one of
// the two roles and the cheaper role's partitions would idle.
static __device__ unsigned g_disk_sm_turn[1024];

template <class S, int C, int W, bool ISMAX, bool TWO, bool HASX>
__global__ void __launch_bounds__(DiskCfg<S, C, W, TWO>::THREADS, DiskCfg<S, C, W, TWO>::MINB)
k_disk(const __grid_constant__ CUtensorMap tm, DiskArgs p, int bw)
{
	using K = DiskCfg<S, C, W, TWO>;
	using D = DiskMarch<S, C, K::LH, ISMAX>;
	constexpr int R = K::R, LH = K::LH, RP = K::RP, PERIOD = K::PERIOD, GP = K::GP, NG = K::NG;
	constexpr int PAIR = K::PAIR, GROUP = K::GROUP;
	extern __shared__ unsigned char smem_raw[];
	// 128-byte aligned base (TMA destination)
	float *ring = reinterpret_cast<float *>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));   // NG groups
	float *tring = ring + (TWO ? (size_t)NG * GROUP : 0);              // the same for the temporaries
	const unsigned bars = smem_u32(tring + (size_t)NG * GROUP);
	const unsigned full_in = bars, empty_in = bars + 8 * NG;
	const unsigned full_t = bars + 16 * NG, empty_t = bars + 24 * NG;

	const int tid = threadIdx.x;
	const int lane = tid & 31;
#ifdef DISK_SWAP
	__shared__ int s_swap;
	if (TWO && tid == 0) {
		unsigned smid;
		asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
		s_swap = (int)(atomicAdd(&g_disk_sm_turn[smid & 1023], 1u) & 1u);
	}
	if (TWO) __syncthreads();
	const bool swap = TWO && s_swap != 0;
#else
	const bool swap = false;
#endif
	const bool first = TWO && ((tid < K::NT) != swap);   // warp-uniform role
	const bool reads_input = !TWO || first;
	const int mt = tid & (K::NT - 1);                  // thread index within its stage
	const bool producer = reads_input && mt == 0;     // issues the TMA copies
	const int plane = blockIdx.z;
	const int cx0 = blockIdx.x * K::OUTW;             // first output column of the strip
	const int o_base = blockIdx.y * p.band_rows;      // first output row of the band (relative)
	const int nout = min(p.band_rows, p.y_rows - o_base);
	const int Y0 = p.y_row0 + o_base;                 // global row of relative output 0
	const int w = p.w, h = p.h;
	const long long pitch = p.pitch;
	const int G2 = (nout + 2 * R + 1) / 2;            // row pairs of the final reduction
	const int Gin = TWO ? G2 + R : G2;                // row pairs of the input ring
	const int Gmine = first ? Gin : G2;
	const int in_row0 = TWO ? Y0 - 2 * R : Y0 - R;    // global row of input pair 0
	const int gc0 = TWO ? cx0 - 2 * LH : cx0 - LH;    // global column of the first window's first float
	// the box starts on a bw boundary at or left of gc0 (floor division, gc0 may be negative)
	const int gxb = (gc0 >= 0 ? gc0 : gc0 - bw + 1) / bw;
	const int shift = gc0 - gxb * bw;                 // 0 <= shift <= 28, multiple of 4

	if (tid == 0) {
		for (int i = 0; i < NG; i++) {
			mbar_init(full_in + 8 * i, 1); mbar_init(empty_in + 8 * i, W);
			mbar_init(full_t + 8 * i, W); mbar_init(empty_t + 8 * i, W);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("prefetch.tensormap [%0];" :: "l"(&tm) : "memory");
	}
	__syncthreads();

	// ---- producer (thread 0): group gi of the input -> ring slot gi mod NG ----
	const unsigned ring_u32 = smem_u32(ring);
	const int trow0 = in_row0 - p.src.row0;           // tensor row of input pair 0
	const int ngroups = (Gin + GP - 1) / GP;
	auto load_group = [&](int gi) {
		const int slot = gi % NG;
		if (gi >= NG) mbar_wait(empty_in + 8 * slot, ((gi / NG) - 1) & 1);
		mbar_arrive_tx(full_in + 8 * slot, K::GROUP_BYTES);
		tma_load_4d(ring_u32 + slot * K::GROUP_BYTES, &tm, 0, gxb, trow0 + 2 * GP * gi, plane, full_in + 8 * slot);
	};
	// Single stage with an epilogue that needs the centre pixel x: the rows of x went through the
	// ring ceil(R/2) pairs before the output rows complete.  When that is less than a group the
	// producer runs one group less ahead, the previous group stays intact while the current one is
	// consumed, and x comes from shared memory instead of a global load whose latency nothing hides
	// (measured: disk7 igradient of the C2 image 0.160 ms with the global load, erosion alone 0.079).
	constexpr bool XRING_OK = !TWO && (R + 1) / 2 + 1 <= GP && NG >= 3;
	const bool xring = XRING_OK && p.xop.p != nullptr && p.xop.p == p.src.p && p.epi != (ISMAX ? EPI_B : EPI_A);
	const int ahead = xring ? NG - 2 : NG - 1;             // groups the producer runs ahead
	if (producer) {
		for (int gi = 0; gi < ahead && gi < ngroups; gi++) load_group(gi);
	}

	// ---- consumer state ----
	float acc[C][K::NACC], hs[C][4];
#pragma unroll
	for (int c = 0; c < C; c++) {
#pragma unroll
		for (int k = 0; k < K::NACC; k++) acc[c][k] = D::init();
#pragma unroll
		for (int k = 0; k < 4; k++) hs[c][k] = D::init();
	}
	int zmin = 0;
	const float *my_ring = reads_input ? ring + shift + C * mt : tring + C * mt;   // this thread's window in slot 0
	float *my_tmp = tring + C * mt;                                     // first stage: where its temporaries go
	const unsigned my_full = reads_input ? full_in : full_t;
	const unsigned my_empty = reads_input ? empty_in : empty_t;
	const bool lane0 = lane == 0;

	// first stage: this thread's temporary columns, global
	const int tx = cx0 - LH + C * mt;
	unsigned tcol_mask = 0;
#pragma unroll
	for (int c = 0; c < C; c++) if (tx + c >= 0 && tx + c < w) tcol_mask |= 1u << c;
	const bool tcol_all = tcol_mask == (1u << C) - 1u;
	// final stage: output columns
	const int x = cx0 + C * mt;
	const bool col_ok = !first && x < w && C * mt < K::OUTW;
	const bool pad_edge = !TWO && col_ok && x + C > w;              // the thread that straddles the true width
	// running pointers to output row 2g-2R of this thread's columns
	float *yq = p.y + plane * p.y_pstride + (long long)(Y0 - p.y_row0 - 2 * R) * pitch + x;
	const float *xq = p.xop.p ? p.xop.p + plane * p.xop.pstride + (long long)(Y0 - p.xop.row0 - 2 * R) * pitch + x : nullptr;
	const float *oq = (!TWO && p.other.p) ? p.other.p + plane * p.other.pstride + (long long)(Y0 - p.other.row0 - 2 * R) * pitch + x : nullptr;
	const int epi = p.epi;
	const bool plain = epi == (ISMAX ? EPI_B : EPI_A);                 // single stage only
	// The one-sided epilogues (src/morsi.c:174,183,252,261) are all y = ec * (ep * x + eq * m) with
	// ep, eq = +-1 and ec = 1 or 0.5: products by +-1 and by 0.5 are exact, so the three roundings are
	// the reference's, and the epilogue is three FMA-pipe instructions instead of an out-of-line switch.
	float ec = 1.f, ep = 1.f, eq = 1.f;
	bool linear = false;
	if (!TWO) {
		if (epi == EPI_X_SUB_A) { linear = true; eq = -1.f; }                       // x - a
		if (epi == EPI_B_SUB_X) { linear = true; ep = -1.f; }                       // b - x
		if (epi == EPI_IBLUR || epi == EPI_EBLUR) { linear = true; ec = 0.5f; }     // (x + m) / 2
	}
	int gi = 0;                                                        // group being consumed
	const float *grp = my_ring;                                        // ... and this thread's window in it
	const float *grp_prev = my_ring;                                   // the group before it (xring)
	unsigned grp_empty = my_empty, grp_empty_prev = my_empty;
	int tgi = 0;                                                       // first stage: temporary group being written
	float *tgrp = my_tmp;
	unsigned tgrp_full = full_t;

#pragma unroll 1
	for (int g0 = 0; g0 < Gmine; g0 += PERIOD) {
#pragma unroll
		for (int s = 0; s < PERIOD; s++) {
			const int g = g0 + s;
			if (g < Gmine) {
				const int pos = s % GP;                                   // place of this pair in its group
				if (pos == 0) {
					// a new group: keep the producer one group ahead, then wait for ours
					const int slot = gi % NG;
					if (producer && gi + ahead < ngroups) load_group(gi + ahead);
					mbar_wait_warp(my_full + 8 * slot, (gi / NG) & 1);
					grp_prev = grp; grp_empty_prev = grp_empty;
					grp = my_ring + slot * GROUP;
					grp_empty = my_empty + 8 * slot;
					gi++;
				}
				const float *rowA = grp + pos * PAIR;
				// the two outputs this step completes
				const int o0 = 2 * g - 2 * R;
				// (one unsigned compare each: 0 <= o < nout)
				const bool e0 = col_ok && (unsigned)o0 < (unsigned)nout;
				const bool e1 = col_ok && (unsigned)(o0 + 1) < (unsigned)nout;
				float xv0[C], xv1[C], ov0[C], ov1[C];
				if (TWO && HASX) {
					if (e0) load_cols<C>(xq, xv0);
					if (e1) load_cols<C>(xq + pitch, xv1);
				}
				if (!TWO && !plain) {
					// operands of the epilogue, in flight during the reductions below
#pragma unroll
					for (int c = 0; c < C; c++) { xv0[c] = 0.f; xv1[c] = 0.f; ov0[c] = 0.f; ov1[c] = 0.f; }
					if (XRING_OK && xring) {
						// output rows o0, o0+1 are input rows 2g-R, 2g-R+1 of the band: pair g - ceil(R/2) (row B
						// of it and row A of the next one when R is odd), at most one group back
						const int back = (R + 1) / 2;
						const int pa = (pos - back + GP) % GP;                  // place of the first row's pair in its group
						const float *ga = pos >= back ? grp : grp_prev;
						const int pb = R % 2 ? (pos - back + 1 + GP) % GP : pa;
						const float *gb = R % 2 ? (pos >= back - 1 ? grp : grp_prev) : ga;
						const float *xa = ga + pa * PAIR + (R % 2 ? RP : 0) + LH;
						const float *xb2 = gb + pb * PAIR + (R % 2 ? 0 : RP) + LH;
#pragma unroll
						for (int c = 0; c < C; c++) { xv0[c] = xa[c]; xv1[c] = xb2[c]; }
					} else {
						if (e0 && xq) load_cols<C>(xq, xv0);
						if (e1 && xq) load_cols<C>(xq + pitch, xv1);
					}
					if (e0 && oq) load_cols<C>(oq, ov0);
					if (e1 && oq) load_cols<C>(oq + pitch, ov1);
				}
				#ifdef DISK_NO_ZCHECK   /* experiment: how much does the -0.0 scan cost the first stage? (results unsafe) */
				D::step(acc, hs, s % (R + 1), rowA, rowA + RP, zmin, false,
#else
				D::step(acc, hs, s % (R + 1), rowA, rowA + RP, zmin, reads_input,
#endif
					// the group has been read (its loads were issued before this
					// arrive and complete long before a refill can land)
					// xring: a group stays in use, as the source of x, for ceil(R/2) steps into the next one
					[&]() {
						if (XRING_OK && xring) { if (pos == (R + 1) / 2 && lane0 && gi >= 2) mbar_arrive(grp_empty_prev); }
						else if (pos == GP - 1 && lane0) mbar_arrive(grp_empty);
					},
					[&](const float (&m0)[C], const float (&m1)[C]) {
					if (first) {
						// temporary pair tp = g-R: rows T0+2tp, T0+2tp+1 (T0 = Y0-R), stored negated
						const int tp = g - R;
						if (tp < 0) return;
						const int tpos = ((s - R) % GP + GP) % GP;
						if (tpos == 0) {
							const int slot = tgi % NG;
							if (tgi >= NG) mbar_wait_warp(empty_t + 8 * slot, ((tgi / NG) - 1) & 1);
							tgrp = my_tmp + slot * GROUP;
							tgrp_full = full_t + 8 * slot;
							tgi++;
						}
						float *q = tgrp + tpos * PAIR;
						const int tr = Y0 - R + 2 * tp;
						float v0[C], v1[C];
#pragma unroll
						for (int c = 0; c < C; c++) { v0[c] = -m0[c]; v1[c] = -m1[c]; }
						if (tcol_all && (unsigned)tr < (unsigned)(h - 1)) {       // rows tr, tr+1 inside the image
							store_cols<C>(q, v0);
							store_cols<C>(q + RP, v1);
						} else {
							disk_store_tmp_masked(q, q + RP, C, tcol_mask, tr >= 0 && tr < h, tr + 1 >= 0 && tr + 1 < h,
								v0[0], v0[1], C == 4 ? v0[2] : 0.f, C == 4 ? v0[3] : 0.f,
								v1[0], v1[1], C == 4 ? v1[2] : 0.f, C == 4 ? v1[3] : 0.f);
						}
						if (tpos == GP - 1 || tp == G2 - 1) {
							__syncwarp();
							if (lane0) mbar_arrive(tgrp_full);
						}
					} else if (TWO) {
						// result = -m (the temporaries were negated)
						if (e0) {
							float v[C];
#pragma unroll
							for (int c = 0; c < C; c++)
								v[c] = !HASX ? -m0[c] : ISMAX ? __fsub_rn(-m0[c], xv0[c]) : __fsub_rn(xv0[c], -m0[c]);
							store_cols<C>(yq, v);
						}
						if (e1) {
							float v[C];
#pragma unroll
							for (int c = 0; c < C; c++)
								v[c] = !HASX ? -m1[c] : ISMAX ? __fsub_rn(-m1[c], xv1[c]) : __fsub_rn(xv1[c], -m1[c]);
							store_cols<C>(yq + pitch, v);
						}
					} else {
						if (e0) {
							if (plain) store_cols<C>(yq, m0);
							else if (linear) {
								float r[C];
#pragma unroll
								for (int c = 0; c < C; c++) r[c] = __fmul_rn(ec, __fadd_rn(__fmul_rn(ep, xv0[c]), __fmul_rn(eq, m0[c])));
								store_cols<C>(yq, r);
							} else {
								const float4 r = disk_epi4<ISMAX>(epi, m0[0], m0[1], C == 4 ? m0[2] : 0.f, C == 4 ? m0[3] : 0.f,
										ov0[0], ov0[1], C == 4 ? ov0[2] : 0.f, C == 4 ? ov0[3] : 0.f,
										xv0[0], xv0[1], C == 4 ? xv0[2] : 0.f, C == 4 ? xv0[3] : 0.f);
								if (C == 4) *(float4 *)yq = r; else *(float2 *)yq = make_float2(r.x, r.y);
							}
						}
						if (e1) {
							if (plain) store_cols<C>(yq + pitch, m1);
							else if (linear) {
								float r[C];
#pragma unroll
								for (int c = 0; c < C; c++) r[c] = __fmul_rn(ec, __fadd_rn(__fmul_rn(ep, xv1[c]), __fmul_rn(eq, m1[c])));
								store_cols<C>(yq + pitch, r);
							} else {
								const float4 r = disk_epi4<ISMAX>(epi, m1[0], m1[1], C == 4 ? m1[2] : 0.f, C == 4 ? m1[3] : 0.f,
										ov1[0], ov1[1], C == 4 ? ov1[2] : 0.f, C == 4 ? ov1[3] : 0.f,
										xv1[0], xv1[1], C == 4 ? xv1[2] : 0.f, C == 4 ? xv1[3] : 0.f);
								if (C == 4) *(float4 *)(yq + pitch) = r; else *(float2 *)(yq + pitch) = make_float2(r.x, r.y);
							}
						}
						// pitched copies (w % 4 != 0): the pad columns of a row stay NaN, the
						// result may be the source of a later pass (oscillation)
						if (pad_edge) {
#pragma unroll
							for (int c = 0; c < C; c++)
								if (x + c >= w) {
									if (e0) yq[c] = CUDART_NAN_F;
									if (e1) yq[pitch + c] = CUDART_NAN_F;
								}
						}
					}
				});
				yq += 2 * pitch;
				if ((TWO && HASX) || !TWO) { if (xq) xq += 2 * pitch; }
				if (!TWO) { if (oq) oq += 2 * pitch; }
			}
		}
	}
	if (__syncthreads_or(zmin == INT_MIN) && tid == 0) atomicOr(p.flag, 1);
}

// ---- erosion AND dilation of one input in one pass ---------------------------------------
// gradient, laplacian, enhance, blur, cblur (src/morsi.c:157-167,187-215,265-275) need the
// erosion a and the dilation b of the SAME image at the same pixel.  Both reductions run
// in one CTA over one TMA ring: warps of role A march the minimum, warps of role B the
// maximum (each group of input rows is handed back once BOTH roles have read it); role A
// leaves its rows in a small shared-memory ring, role B picks them up when the same output
// rows complete on its side, applies the epilogue with the centre pixel x (re-read from
// global memory, an L2 hit: the rows went through the ring R rows earlier) and stores.
// 8 B/sample of HBM traffic instead of the 16-24 B of two passes through a temporary image.
template <class S, int C, int W>
struct DiskCfgBoth {
	using K1 = DiskCfg<S, C, W, false>;
	static constexpr int R = K1::R, LH = K1::LH, NT = K1::NT, TW = K1::TW, OUTW = K1::OUTW, RP = K1::RP;
	static constexpr int PERIOD = K1::PERIOD, GP = K1::GP, PAIR = K1::PAIR, GROUP = K1::GROUP, NG = 2;
	static constexpr unsigned GROUP_BYTES = K1::GROUP_BYTES;
	static constexpr int APAIR = 2 * TW;              // role A's rows: just the strip's own columns
	static constexpr int AGROUP = GP * APAIR;
	static constexpr int NACC = K1::NACC;
	static constexpr int THREADS = 2 * NT;
	static constexpr size_t SMEM = (size_t)NG * (GROUP + AGROUP) * sizeof(float) + 4 * NG * sizeof(unsigned long long) + 128;
	static constexpr int REGS = (C * (2 * R + 2) > 64) ? 255 : (THREADS >= 256 ? DISK_REGS_SMALL_W4 : DISK_REGS_SMALL);
	static constexpr int MINB = 65536 / (REGS * THREADS) > 0 ? 65536 / (REGS * THREADS) : 1;
};

template <class S, int C, int W>
__global__ void __launch_bounds__(DiskCfgBoth<S, C, W>::THREADS, DiskCfgBoth<S, C, W>::MINB)
k_disk_both(const __grid_constant__ CUtensorMap tm, DiskArgs p, int bw)
{
	using K = DiskCfgBoth<S, C, W>;
	constexpr int R = K::R, LH = K::LH, RP = K::RP, PERIOD = K::PERIOD, GP = K::GP, NG = K::NG;
	constexpr int PAIR = K::PAIR, GROUP = K::GROUP, APAIR = K::APAIR, AGROUP = K::AGROUP, TW = K::TW;
	extern __shared__ unsigned char smem_raw[];
	float *ring = reinterpret_cast<float *>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));   // NG groups of input rows
	float *aring = ring + (size_t)NG * GROUP;                          // NG groups of role A's result rows
	const unsigned bars = smem_u32(aring + (size_t)NG * AGROUP);
	const unsigned full_in = bars, empty_in = bars + 8 * NG;
	const unsigned full_a = bars + 16 * NG, empty_a = bars + 24 * NG;
	__shared__ int s_swap;

	const int tid = threadIdx.x;
	const int lane = tid & 31;
	if (tid == 0) {
		for (int i = 0; i < NG; i++) {
			mbar_init(full_in + 8 * i, 1); mbar_init(empty_in + 8 * i, 2 * W);
			mbar_init(full_a + 8 * i, W); mbar_init(empty_a + 8 * i, W);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("prefetch.tensormap [%0];" :: "l"(&tm) : "memory");
		unsigned smid;
		asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
		s_swap = (int)(atomicAdd(&g_disk_sm_turn[smid & 1023], 1u) & 1u);
	}
	__syncthreads();
	const bool first = (tid < K::NT) != (s_swap != 0);   // role A (minimum); warp-uniform
	const int mt = tid & (K::NT - 1);
	const bool producer = first && mt == 0;
	const int plane = blockIdx.z;
	const int cx0 = blockIdx.x * K::OUTW;
	const int o_base = blockIdx.y * p.band_rows;
	const int nout = min(p.band_rows, p.y_rows - o_base);
	const int Y0 = p.y_row0 + o_base;
	const int w = p.w;
	const long long pitch = p.pitch;
	const int G2 = (nout + 2 * R + 1) / 2;            // row pairs marched
	const int NOP = (nout + 1) / 2;                   // output row pairs
	const int in_row0 = Y0 - R;
	const int gc0 = cx0 - LH;
	const int gxb = (gc0 >= 0 ? gc0 : gc0 - bw + 1) / bw;
	const int shift = gc0 - gxb * bw;

	const unsigned ring_u32 = smem_u32(ring);
	const int trow0 = in_row0 - p.src.row0;
	const int ngroups = (G2 + GP - 1) / GP;
	// Role B holds input groups too, and it may sit in the middle of one while it waits for
	// rows of role A: the producer (a thread of role A) must therefore never BLOCK on a
	// slot while role A still has rows to deliver from the groups it already holds.  It
	// tries at every step, without waiting, and insists only at the start of the group it
	// needs itself -- by then role A has delivered everything role B can be waiting for.
	int next_load = 0;
	auto issue_group = [&](int gl) {
		const int slot = gl % NG;
		mbar_arrive_tx(full_in + 8 * slot, K::GROUP_BYTES);
		tma_load_4d(ring_u32 + slot * K::GROUP_BYTES, &tm, 0, gxb, trow0 + 2 * GP * gl, plane, full_in + 8 * slot);
	};
	if (producer) {
		for (; next_load < NG - 1 && next_load < ngroups; next_load++) issue_group(next_load);
	}

	const float *my_ring = ring + shift + C * mt;
	float *my_a = aring + C * mt;
	const bool lane0 = lane == 0;
	const int x = cx0 + C * mt;
	const bool col_ok = !first && x < w;
	const bool pad_edge = col_ok && x + C > w;
	float *yq = p.y + plane * p.y_pstride + (long long)(Y0 - p.y_row0 - 2 * R) * pitch + x;
	const float *xq = p.xop.p ? p.xop.p + plane * p.xop.pstride + (long long)(Y0 - p.xop.row0 - 2 * R) * pitch + x : nullptr;
	const int epi = p.epi;
	int zmin = 0;

	auto run = [&](auto flavour) {
		constexpr bool ISMAX = decltype(flavour)::value;
		using D = DiskMarch<S, C, LH, ISMAX>;
		float acc[C][K::NACC], hs[C][4];
#pragma unroll
		for (int c = 0; c < C; c++) {
#pragma unroll
			for (int k = 0; k < K::NACC; k++) acc[c][k] = D::init();
#pragma unroll
			for (int k = 0; k < 4; k++) hs[c][k] = D::init();
		}
		int gi = 0, agi = 0;
		const float *grp = my_ring;
		float *agrp = my_a;
		unsigned grp_empty = empty_in, agrp_bar = full_a;
#pragma unroll 1
		for (int g0 = 0; g0 < G2; g0 += PERIOD) {
#pragma unroll
			for (int s = 0; s < PERIOD; s++) {
				const int g = g0 + s;
				if (g < G2) {
					const int pos = s % GP;
					if (producer) {
						// gi - (pos != 0) = the group being consumed; group n may go into its slot once group n - NG is done
						const int cur = pos == 0 ? gi : gi - 1;
						if (pos == 0)
							while (next_load <= cur && next_load < ngroups) {           // needed now: wait for the slot
								if (next_load >= NG) mbar_wait(empty_in + 8 * (next_load % NG), ((next_load / NG) - 1) & 1);
								issue_group(next_load++);
							}
						if (next_load < ngroups && next_load < cur + NG &&
								(next_load < NG || mbar_try(empty_in + 8 * (next_load % NG), ((next_load / NG) - 1) & 1)))
							issue_group(next_load++);
					}
					if (pos == 0) {
						const int slot = gi % NG;
						mbar_wait_warp(full_in + 8 * slot, (gi / NG) & 1);
						grp = my_ring + slot * GROUP;
						grp_empty = empty_in + 8 * slot;
						gi++;
					}
					const float *rowA = grp + pos * PAIR;
					const int o0 = 2 * g - 2 * R;
					const bool e0 = col_ok && (unsigned)o0 < (unsigned)nout;
					const bool e1 = col_ok && (unsigned)(o0 + 1) < (unsigned)nout;
					float xv0[C], xv1[C];
					if (ISMAX) {
						// role B: the centre pixels of the two output rows, in flight during the reductions
#pragma unroll
						for (int c = 0; c < C; c++) { xv0[c] = 0.f; xv1[c] = 0.f; }
						if (e0 && xq) load_cols<C>(xq, xv0);
						if (e1 && xq) load_cols<C>(xq + pitch, xv1);
					}
					D::step(acc, hs, s % (R + 1), rowA, rowA + RP, zmin, !ISMAX,
						[&]() { if (pos == GP - 1 && lane0) mbar_arrive(grp_empty); },
						[&](const float (&m0)[C], const float (&m1)[C]) {
						const int op = g - R;                                   // output row pair
						if (op < 0) return;
						const int apos = ((s - R) % GP + GP) % GP;
						if (!ISMAX) {
							// role A: rows 2 op, 2 op + 1 of the erosion -> ring
							if (apos == 0) {
								const int slot = agi % NG;
								if (agi >= NG) mbar_wait_warp(empty_a + 8 * slot, ((agi / NG) - 1) & 1);
								agrp = my_a + slot * AGROUP;
								agrp_bar = full_a + 8 * slot;
								agi++;
							}
							float *q = agrp + apos * APAIR;
							store_cols<C>(q, m0);
							store_cols<C>(q + TW, m1);
							if (apos == GP - 1 || op == NOP - 1) {
								__syncwarp();
								if (lane0) mbar_arrive(agrp_bar);
							}
						} else {
							// role B: the dilation is in m0 / m1; fetch the erosion, finish, store
							if (apos == 0) {
								const int slot = agi % NG;
								mbar_wait_warp(full_a + 8 * slot, (agi / NG) & 1);
								agrp = my_a + slot * AGROUP;
								agrp_bar = empty_a + 8 * slot;
								agi++;
							}
							const float *q = agrp + apos * APAIR;
							float a0[C], a1[C];
							if (C == 4) {
								const float4 t0 = *(const float4 *)q, t1 = *(const float4 *)(q + TW);
								a0[0] = t0.x; a0[1] = t0.y; a0[2] = t0.z; a0[3] = t0.w;
								a1[0] = t1.x; a1[1] = t1.y; a1[2] = t1.z; a1[3] = t1.w;
							} else {
								const float2 t0 = *(const float2 *)q, t1 = *(const float2 *)(q + TW);
								a0[0] = t0.x; a0[1] = t0.y; a1[0] = t1.x; a1[1] = t1.y;
							}
							if (e0) {
								const float4 r = disk_epi4<true>(epi, m0[0], m0[1], C == 4 ? m0[2] : 0.f, C == 4 ? m0[3] : 0.f,
										a0[0], a0[1], C == 4 ? a0[2] : 0.f, C == 4 ? a0[3] : 0.f,
										xv0[0], xv0[1], C == 4 ? xv0[2] : 0.f, C == 4 ? xv0[3] : 0.f);
								if (C == 4) *(float4 *)yq = r; else *(float2 *)yq = make_float2(r.x, r.y);
							}
							if (e1) {
								const float4 r = disk_epi4<true>(epi, m1[0], m1[1], C == 4 ? m1[2] : 0.f, C == 4 ? m1[3] : 0.f,
										a1[0], a1[1], C == 4 ? a1[2] : 0.f, C == 4 ? a1[3] : 0.f,
										xv1[0], xv1[1], C == 4 ? xv1[2] : 0.f, C == 4 ? xv1[3] : 0.f);
								if (C == 4) *(float4 *)(yq + pitch) = r; else *(float2 *)(yq + pitch) = make_float2(r.x, r.y);
							}
							// the values are in registers (the epilogue consumed them): hand the group back
							if (apos == GP - 1 || op == NOP - 1) {
								__syncwarp();
								if (lane0) mbar_arrive(agrp_bar);
							}
							if (pad_edge) {
#pragma unroll
								for (int c = 0; c < C; c++)
									if (x + c >= w) {
										if (e0) yq[c] = CUDART_NAN_F;
										if (e1) yq[pitch + c] = CUDART_NAN_F;
									}
							}
						}
					});
					if (ISMAX) {
						yq += 2 * pitch;
						if (xq) xq += 2 * pitch;
					}
				}
			}
		}
	};
	if (first) run(std::false_type{});
	else run(std::true_type{});
	if (__syncthreads_or(zmin == INT_MIN) && tid == 0) atomicOr(p.flag, 1);
}

// ---- the same, with both reductions in EVERY thread (small disks) -------------------------
// For R <= 6 a thread can hold the accumulators of the minimum AND of the maximum
// (2 x C x (2R+2) registers): one role, one ring, no hand-off between warps -- the two
// marches read the same shared-memory rows and are independent instruction streams that
// fill each other's latency.  Epilogue and store as in k_disk_both.
template <class S, int C, int W>
struct DiskCfgDual {
	using K1 = DiskCfg<S, C, W, false>;
	static constexpr int R = K1::R, LH = K1::LH, NT = K1::NT, TW = K1::TW, OUTW = K1::OUTW, RP = K1::RP;
	static constexpr int PERIOD = K1::PERIOD, GP = K1::GP, PAIR = K1::PAIR, GROUP = K1::GROUP, NG = 3;
	static constexpr unsigned GROUP_BYTES = K1::GROUP_BYTES;
	static constexpr int NACC = K1::NACC;
	static constexpr int THREADS = NT;
	static constexpr size_t SMEM = (size_t)NG * GROUP * sizeof(float) + 2 * NG * sizeof(unsigned long long) + 128;
	static constexpr int MINB = 65536 / (255 * THREADS) > 0 ? 65536 / (255 * THREADS) : 1;
	static constexpr bool OK = 2 * C * (2 * R + 2) <= 112;     // the accumulators of both flavours fit the register file
};

// EPI: the epilogue, compile time (a run-time switch in an out-of-line call cost 15 % of the samples here)
template <class S, int C, int W, int EPI>
__global__ void __launch_bounds__(DiskCfgDual<S, C, W>::THREADS, DiskCfgDual<S, C, W>::MINB)
k_disk_dual(const __grid_constant__ CUtensorMap tm, DiskArgs p, int bw)
{
	using K = DiskCfgDual<S, C, W>;
	using Dmin = DiskMarch<S, C, K::LH, false>;
	using Dmax = DiskMarch<S, C, K::LH, true>;
	constexpr int R = K::R, LH = K::LH, RP = K::RP, PERIOD = K::PERIOD, GP = K::GP, NG = K::NG;
	constexpr int PAIR = K::PAIR, GROUP = K::GROUP;
	extern __shared__ unsigned char smem_raw[];
	float *ring = reinterpret_cast<float *>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
	const unsigned bars = smem_u32(ring + (size_t)NG * GROUP);
	const unsigned full_in = bars, empty_in = bars + 8 * NG;

	const int tid = threadIdx.x;
	const int lane = tid & 31;
	const int plane = blockIdx.z;
	const int cx0 = blockIdx.x * K::OUTW;
	const int o_base = blockIdx.y * p.band_rows;
	const int nout = min(p.band_rows, p.y_rows - o_base);
	const int Y0 = p.y_row0 + o_base;
	const int w = p.w;
	const long long pitch = p.pitch;
	const int G2 = (nout + 2 * R + 1) / 2;
	const int in_row0 = Y0 - R;
	const int gc0 = cx0 - LH;
	const int gxb = (gc0 >= 0 ? gc0 : gc0 - bw + 1) / bw;
	const int shift = gc0 - gxb * bw;

	if (tid == 0) {
		for (int i = 0; i < NG; i++) { mbar_init(full_in + 8 * i, 1); mbar_init(empty_in + 8 * i, W); }
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("prefetch.tensormap [%0];" :: "l"(&tm) : "memory");
	}
	__syncthreads();

	const unsigned ring_u32 = smem_u32(ring);
	const int trow0 = in_row0 - p.src.row0;
	const int ngroups = (G2 + GP - 1) / GP;
	auto load_group = [&](int gl) {
		const int slot = gl % NG;
		if (gl >= NG) mbar_wait(empty_in + 8 * slot, ((gl / NG) - 1) & 1);
		mbar_arrive_tx(full_in + 8 * slot, K::GROUP_BYTES);
		tma_load_4d(ring_u32 + slot * K::GROUP_BYTES, &tm, 0, gxb, trow0 + 2 * GP * gl, plane, full_in + 8 * slot);
	};
	// epilogues with the centre pixel take it from the ring (see k_disk): the producer then runs one group ahead
	constexpr bool XRING = EpiNeeds<EPI>::x && (R + 1) / 2 + 1 <= GP;
	constexpr int AHEAD = XRING ? NG - 2 : NG - 1;
	if (tid == 0) {
		for (int gl = 0; gl < AHEAD && gl < ngroups; gl++) load_group(gl);
	}

	float accA[C][K::NACC], hsA[C][4], accB[C][K::NACC], hsB[C][4];
#pragma unroll
	for (int c = 0; c < C; c++) {
#pragma unroll
		for (int k = 0; k < K::NACC; k++) { accA[c][k] = Dmin::init(); accB[c][k] = Dmax::init(); }
#pragma unroll
		for (int k = 0; k < 4; k++) { hsA[c][k] = Dmin::init(); hsB[c][k] = Dmax::init(); }
	}
	int zmin = 0, zdummy = 0;
	const float *my_ring = ring + shift + C * tid;
	const bool lane0 = lane == 0;
	const int x = cx0 + C * tid;
	const bool col_ok = x < w;
	const bool pad_edge = col_ok && x + C > w;
	float *yq = p.y + plane * p.y_pstride + (long long)(Y0 - p.y_row0 - 2 * R) * pitch + x;
	const float *xq = p.xop.p ? p.xop.p + plane * p.xop.pstride + (long long)(Y0 - p.xop.row0 - 2 * R) * pitch + x : nullptr;
	int gi = 0;
	const float *grp = my_ring, *grp_prev = my_ring;
	unsigned grp_empty = empty_in, grp_empty_prev = empty_in;

#pragma unroll 1
	for (int g0 = 0; g0 < G2; g0 += PERIOD) {
#pragma unroll
		for (int s = 0; s < PERIOD; s++) {
			const int g = g0 + s;
			if (g < G2) {
				const int pos = s % GP;
				if (pos == 0) {
					const int slot = gi % NG;
					if (tid == 0 && gi + AHEAD < ngroups) load_group(gi + AHEAD);
					mbar_wait_warp(full_in + 8 * slot, (gi / NG) & 1);
					grp_prev = grp; grp_empty_prev = grp_empty;
					grp = my_ring + slot * GROUP;
					grp_empty = empty_in + 8 * slot;
					gi++;
				}
				const float *rowA = grp + pos * PAIR;
				const int o0 = 2 * g - 2 * R;
				const bool e0 = col_ok && (unsigned)o0 < (unsigned)nout;
				const bool e1 = col_ok && (unsigned)(o0 + 1) < (unsigned)nout;
				float xv0[C], xv1[C], a0[C], a1[C];
#pragma unroll
				for (int c = 0; c < C; c++) { xv0[c] = 0.f; xv1[c] = 0.f; }
				if (XRING) {
					const int back = (R + 1) / 2;
					const int pa = (pos - back + GP) % GP;
					const float *ga = pos >= back ? grp : grp_prev;
					const int pb = R % 2 ? (pos - back + 1 + GP) % GP : pa;
					const float *gb = R % 2 ? (pos >= back - 1 ? grp : grp_prev) : ga;
					const float *xa = ga + pa * PAIR + (R % 2 ? RP : 0) + LH;
					const float *xb2 = gb + pb * PAIR + (R % 2 ? 0 : RP) + LH;
#pragma unroll
					for (int c = 0; c < C; c++) { xv0[c] = xa[c]; xv1[c] = xb2[c]; }
				} else if (EpiNeeds<EPI>::x) {
					if (e0) load_cols<C>(xq, xv0);
					if (e1) load_cols<C>(xq + pitch, xv1);
				}
				Dmin::step(accA, hsA, s % (R + 1), rowA, rowA + RP, zmin, true,
					[]() {},
					[&](const float (&m0)[C], const float (&m1)[C]) {
#pragma unroll
						for (int c = 0; c < C; c++) { a0[c] = m0[c]; a1[c] = m1[c]; }
					});
				Dmax::step(accB, hsB, s % (R + 1), rowA, rowA + RP, zdummy, false,
					[&]() {
						if (XRING) { if (pos == (R + 1) / 2 && lane0 && gi >= 2) mbar_arrive(grp_empty_prev); }
						else if (pos == GP - 1 && lane0) mbar_arrive(grp_empty);
					},
					[&](const float (&m0)[C], const float (&m1)[C]) {
					if (e0) {
						float r[C];
#pragma unroll
						for (int c = 0; c < C; c++) r[c] = epilogue<EPI>(a0[c], m0[c], xv0[c]);
						store_cols<C>(yq, r);
					}
					if (e1) {
						float r[C];
#pragma unroll
						for (int c = 0; c < C; c++) r[c] = epilogue<EPI>(a1[c], m1[c], xv1[c]);
						store_cols<C>(yq + pitch, r);
					}
					if (pad_edge) {
#pragma unroll
						for (int c = 0; c < C; c++)
							if (x + c >= w) {
								if (e0) yq[c] = CUDART_NAN_F;
								if (e1) yq[pitch + c] = CUDART_NAN_F;
							}
					}
				});
				yq += 2 * pitch;
				if (xq) xq += 2 * pitch;
			}
		}
	}
	if (__syncthreads_or(zmin == INT_MIN) && tid == 0) atomicOr(p.flag, 1);
}

// ---- host side --------------------------------------------------------------------
// Bands: every CTA marches `rows` output rows plus a warm-up of 2*reach rows
// per stage, and CTAs run in waves of `slots`; pick the band count that
// minimises waves x (rows + warm-up), i.e. no half-empty last wave.
static int pick_band_rows(long long slots, int y_rows, long long strips_x_planes, int reach, int stages, double *cost_out)
{
	// rows of warm-up + fixed per-CTA cost.  The two stages of a fused CTA warm
	// up side by side (different warps), so the measured cost is one stage's:
	// C2 on B200 (51 strip-planes, 444 CTA slots): 8 bands 0.324 ms, 17 bands
	// 0.299 ms, 26 bands 0.298 ms, 9 bands 0.373 ms (a barely started 2nd wave).
	(void)stages;
	const int warm = 2 * reach + 12;
	const int min_rows = 8 * reach * stages > 32 ? 8 * reach * stages : 32;
	int best_rows = y_rows;
	double best = 1e300;
	static const int forced_bands = getenv("MORSI_DISK_BANDS") ? atoi(getenv("MORSI_DISK_BANDS")) : 0;   // experiments
	if (forced_bands > 0) {
		int rows = (y_rows + forced_bands - 1) / forced_bands;
		rows = (rows + 1) & ~1;
		if (cost_out) *cost_out = 1.0;
		return rows;
	}
	for (int bands = 1; bands <= 4096; bands++) {
		int rows = (y_rows + bands - 1) / bands;
		rows = (rows + 1) & ~1;
		if (rows < min_rows && bands > 1) break;
		const long long ctas = strips_x_planes * ((y_rows + rows - 1) / rows);
		const long long waves = (ctas + slots - 1) / slots;
		const double cost = (double)waves * (rows + warm);
		if (cost < best * 0.999) { best = cost; best_rows = rows; }
	}
	if (cost_out) *cost_out = best;
	return best_rows;
}

template <class S, int C, int W, bool ISMAX, bool TWO, bool HASX>
static int disk_occupancy(int device)
{
	using K = DiskCfg<S, C, W, TWO>;
	static int occs[64];                               // per device: 0 = not queried yet
	int &occ = occs[device & 63];
	if (occ <= 0) {
		cudaFuncSetAttribute(k_disk<S, C, W, ISMAX, TWO, HASX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM);
		int o = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_disk<S, C, W, ISMAX, TWO, HASX>, K::THREADS, K::SMEM) != cudaSuccess || o < 1)
			o = 1;
		occ = o;
	}
	return occ;
}

// cost (in marched rows per SM slot, scaled by the warps a CTA keeps busy) of
// running the job with W warps per stage
template <class S, int C, int W, bool ISMAX, bool TWO, bool HASX>
static double disk_plan(const MorsiCtx *c, const DiskArgs &a, int planes, int *band_rows)
{
	using K = DiskCfg<S, C, W, TWO>;
	const int strips = (a.w + K::OUTW - 1) / K::OUTW;
	const int occ = disk_occupancy<S, C, W, ISMAX, TWO, HASX>(c->device);
	double cost;
	*band_rows = pick_band_rows((long long)c->sm_count * occ, a.y_rows, (long long)strips * planes, S::R, TWO ? 2 : 1, &cost);
	// a CTA's speed is proportional to 1 / (warps resident on its SM)
	return cost * occ * W;
}

// The source band as a 4-D tensor (bw, w/bw, rows, planes): dimension 0 is a
// run of bw contiguous floats, dimension 1 steps over such runs, so a box
// (bw, RP/bw, nrows, 1) is nrows dense rows of RP floats in shared memory
// whatever RP is (a plain 2-D box is limited to 256 columns).  Everything
// outside [0,w) x [0,rows) reads as NaN.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
		const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
		CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int disk_tensor_map(CUtensorMap *tm, const Band &src, int src_rows, int w /* = the row pitch */, int planes, int rp, int nrows, int *bw_out)
{
	static EncodeTiledFn encode = nullptr;
	if (!encode) {
		void *fn = nullptr;
		cudaDriverEntryPointQueryResult qres;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
			return morsi_set_error(MORSI_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
		encode = (EncodeTiledFn)fn;
	}
	int bw = 32;
	while (w % bw) bw >>= 1;                           // w % 4 == 0 is a precondition of this path
	const long long ps = planes > 1 ? src.pstride : (long long)w * src_rows;
	cuuint64_t dims[4] = {(cuuint64_t)bw, (cuuint64_t)(w / bw), (cuuint64_t)src_rows, (cuuint64_t)planes};
	cuuint64_t strides[3] = {(cuuint64_t)bw * 4, (cuuint64_t)w * 4, (cuuint64_t)ps * 4};
	cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)(rp / bw), (cuuint32_t)nrows, 1};
	cuuint32_t estr[4] = {1, 1, 1, 1};
	CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)src.p, dims, strides, box, estr,
			CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
			CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
	if (r != CUDA_SUCCESS)
		return morsi_set_error(MORSI_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for %dx%dx%d, pitch %lld", (int)r, w, src_rows, planes, ps);
	*bw_out = bw;
	return MORSI_OK;
}

template <class S, int C, int W, bool ISMAX, bool TWO, bool HASX>
static int disk_launch(const MorsiCtx *c, const DiskArgs &a0, int planes, int band_rows, cudaStream_t st)
{
	using K = DiskCfg<S, C, W, TWO>;
	DiskArgs a = a0;
	a.band_rows = band_rows;
	CUtensorMap tm;
	int bw = 4;
	// over the row PITCH: pad columns [w, pitch) hold NaN (absent, like everything outside)
	int rc = disk_tensor_map(&tm, a.src, a.src_rows, a.pitch, planes, K::RP, 2 * K::GP, &bw);
	if (rc) return rc;
	const int strips = (a.w + K::OUTW - 1) / K::OUTW;
	dim3 grid(strips, (a.y_rows + band_rows - 1) / band_rows, planes);
	disk_occupancy<S, C, W, ISMAX, TWO, HASX>(c->device);    // sets the shared-memory attribute once
	k_disk<S, C, W, ISMAX, TWO, HASX><<<grid, K::THREADS, K::SMEM, st>>>(tm, a, bw);
	morsi_count_launch(1);
	MORSI_CU(cudaGetLastError());
	return MORSI_OK;
}

template <class S, int C, int W>
static int disk_both_occupancy(int device)
{
	using K = DiskCfgBoth<S, C, W>;
	static int occs[64];
	int &occ = occs[device & 63];
	if (occ <= 0) {
		cudaFuncSetAttribute(k_disk_both<S, C, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM);
		int o = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_disk_both<S, C, W>, K::THREADS, K::SMEM) != cudaSuccess || o < 1)
			o = 1;
		occ = o;
	}
	return occ;
}

template <class S, int C, int W>
static int disk_launch_both(const MorsiCtx *c, const DiskArgs &a0, int planes, cudaStream_t st)
{
	using K = DiskCfgBoth<S, C, W>;
	DiskArgs a = a0;
	const int strips = (a.w + K::OUTW - 1) / K::OUTW;
	const int occ = disk_both_occupancy<S, C, W>(c->device);
	a.band_rows = pick_band_rows((long long)c->sm_count * occ, a.y_rows, (long long)strips * planes, S::R, 1, nullptr);
	CUtensorMap tm;
	int bw = 4;
	int rc = disk_tensor_map(&tm, a.src, a.src_rows, a.pitch, planes, K::RP, 2 * K::GP, &bw);
	if (rc) return rc;
	dim3 grid(strips, (a.y_rows + a.band_rows - 1) / a.band_rows, planes);
	k_disk_both<S, C, W><<<grid, K::THREADS, K::SMEM, st>>>(tm, a, bw);
	morsi_count_launch(1);
	MORSI_CU(cudaGetLastError());
	return MORSI_OK;
}

template <class S, int C, int W, int EPI>
static int disk_launch_dual_e(const MorsiCtx *c, const DiskArgs &a0, int planes, cudaStream_t st)
{
	using K = DiskCfgDual<S, C, W>;
	DiskArgs a = a0;
	static int occs[64];
	int &occ = occs[c->device & 63];
	if (occ <= 0) {
		cudaFuncSetAttribute(k_disk_dual<S, C, W, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM);
		int o = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_disk_dual<S, C, W, EPI>, K::THREADS, K::SMEM) != cudaSuccess || o < 1) o = 1;
		occ = o;
	}
	const int strips = (a.w + K::OUTW - 1) / K::OUTW;
	a.band_rows = pick_band_rows((long long)c->sm_count * occ, a.y_rows, (long long)strips * planes, S::R, 1, nullptr);
	CUtensorMap tm;
	int bw = 4;
	int rc = disk_tensor_map(&tm, a.src, a.src_rows, a.pitch, planes, K::RP, 2 * K::GP, &bw);
	if (rc) return rc;
	dim3 grid(strips, (a.y_rows + a.band_rows - 1) / a.band_rows, planes);
	k_disk_dual<S, C, W, EPI><<<grid, K::THREADS, K::SMEM, st>>>(tm, a, bw);
	morsi_count_launch(1);
	MORSI_CU(cudaGetLastError());
	return MORSI_OK;
}

// -1: not an epilogue of the gradient family (the caller takes another kernel)
template <class S, int C, int W>
static int disk_launch_dual(const MorsiCtx *c, const DiskArgs &a, int planes, cudaStream_t st)
{
	switch (a.epi) {
	case EPI_B_SUB_A: return disk_launch_dual_e<S, C, W, EPI_B_SUB_A>(c, a, planes, st);
	case EPI_LAP: return disk_launch_dual_e<S, C, W, EPI_LAP>(c, a, planes, st);
	case EPI_ENH: return disk_launch_dual_e<S, C, W, EPI_ENH>(c, a, planes, st);
	case EPI_BLUR: return disk_launch_dual_e<S, C, W, EPI_BLUR>(c, a, planes, st);
	case EPI_CBLUR: return disk_launch_dual_e<S, C, W, EPI_CBLUR>(c, a, planes, st);
	}
	return -1;
}

static int disk_forced_w()
{
	const char *s = getenv("MORSI_DISK_W");            // 2 or 4 warps per stage; default: planned
	return s ? atoi(s) : 0;
}

template <class S, int C, bool ISMAX, bool TWO, bool HASX>
static int disk_run(MorsiCtx *c, const DiskArgs &a, int planes, cudaStream_t st)
{
	int rows2 = 0, rows4 = 0;
	const double c2 = disk_plan<S, C, 2, ISMAX, TWO, HASX>(c, a, planes, &rows2);
	const double c4 = disk_plan<S, C, 4, ISMAX, TWO, HASX>(c, a, planes, &rows4);
	const int force = disk_forced_w();
	// measured on B200 (C2, disk7): the 4-warp-per-stage CTA never beats the 2-warp one
	// for the small disks even when the model calls it even; the big disks (255
	// registers, one CTA per SM either way) prefer 4 (C4: 8.5 ms vs 10.3 ms)
	const bool small = DiskCfg<S, C, 2, TWO>::REGS < 255;
	if (force == 2 || (force != 4 && c2 < c4 * (small ? 1.3 : 1.0)))
		return disk_launch<S, C, 2, ISMAX, TWO, HASX>(c, a, planes, rows2, st);
	return disk_launch<S, C, 4, ISMAX, TWO, HASX>(c, a, planes, rows4, st);
}

template <int ID, int C>
static int disk_shape_c(MorsiCtx *c, const DiskArgs &a, int planes, bool ismax, bool two, cudaStream_t st)
{
	using S = Shape<ID>;
	if (a.both) {
		// both reductions in every thread where the register file holds them (R <= 6; MORSI_DISK_DUAL=0: the two-role kernel)
		static const bool no_dual = getenv("MORSI_DISK_DUAL") && !strcmp(getenv("MORSI_DISK_DUAL"), "0");
		if (DiskCfgDual<S, C, 2>::OK && !no_dual) {
			if constexpr (DiskCfgDual<S, C, 2>::OK) {
				const int rc = disk_launch_dual<S, C, 2>(c, a, planes, st);
				if (rc != -1) return rc;
			}
		}
		// 2 warps per role for the small disks, 4 for the ones that need every register
		if (disk_forced_w() == 4 || (disk_forced_w() != 2 && DiskCfgBoth<S, C, 2>::REGS >= 255))
			return disk_launch_both<S, C, 4>(c, a, planes, st);
		return disk_launch_both<S, C, 2>(c, a, planes, st);
	}
	if (two && a.xop.p)
		return ismax ? disk_run<S, C, true, true, true>(c, a, planes, st) : disk_run<S, C, false, true, true>(c, a, planes, st);
	if (two)
		return ismax ? disk_run<S, C, true, true, false>(c, a, planes, st) : disk_run<S, C, false, true, false>(c, a, planes, st);
	return ismax ? disk_run<S, C, true, false, false>(c, a, planes, st) : disk_run<S, C, false, false, false>(c, a, planes, st);
}

template <int ID>
static int disk_shape(MorsiCtx *c, const DiskArgs &a, int planes, bool ismax, bool two, cudaStream_t st)
{
#ifdef MORSI_DISK_EXPERIMENT_C2
	static const bool c2 = getenv("MORSI_DISK_C") && atoi(getenv("MORSI_DISK_C")) == 2;
	if (c2) return disk_shape_c<ID, 2>(c, a, planes, ismax, two, st);
#endif
	return disk_shape_c<ID, 4>(c, a, planes, ismax, two, st);
}

// The shapes are compiled in four translation units -- this source with
// -DMORSI_DISK_PART=0..3 (Makefile) -- so that the build parallelises; part 0
// also holds the host logic.  Development aid: -DMORSI_DISK_IDS="T(8) T(13)"
// compiles a subset of the shapes, all in part 0.
#ifndef MORSI_DISK_PART
#define MORSI_DISK_PART 0
#endif
#ifdef MORSI_DISK_IDS
#define DISK_IDS_P0 MORSI_DISK_IDS
#define DISK_IDS_P1
#define DISK_IDS_P2
#define DISK_IDS_P3
#else
#define DISK_IDS_P0 T(0) T(1) T(2) T(3) T(4) T(5) T(6) T(7)
#define DISK_IDS_P1 T(8) T(9) T(10) T(11)
#define DISK_IDS_P2 T(12) T(16) T(17)
#define DISK_IDS_P3 T(13) T(18)
#endif
#define DISK_IDS DISK_IDS_P0 DISK_IDS_P1 DISK_IDS_P2 DISK_IDS_P3
#if MORSI_DISK_PART == 0
#define DISK_IDS_MINE DISK_IDS_P0
#elif MORSI_DISK_PART == 1
#define DISK_IDS_MINE DISK_IDS_P1
#elif MORSI_DISK_PART == 2
#define DISK_IDS_MINE DISK_IDS_P2
#else
#define DISK_IDS_MINE DISK_IDS_P3
#endif

// launch shape `id` if this part compiled it; -2: not mine
#define DISK_PART_FN2(N) morsi_disk_launch_part##N
#define DISK_PART_FN(N) DISK_PART_FN2(N)
int morsi_disk_launch_part0(int id, MorsiCtx *c, const DiskArgs &a, int planes, bool ismax, bool two, cudaStream_t st);
int morsi_disk_launch_part1(int id, MorsiCtx *c, const DiskArgs &a, int planes, bool ismax, bool two, cudaStream_t st);
int morsi_disk_launch_part2(int id, MorsiCtx *c, const DiskArgs &a, int planes, bool ismax, bool two, cudaStream_t st);
int morsi_disk_launch_part3(int id, MorsiCtx *c, const DiskArgs &a, int planes, bool ismax, bool two, cudaStream_t st);

int DISK_PART_FN(MORSI_DISK_PART)(int id, MorsiCtx *c, const DiskArgs &a, int planes, bool ismax, bool two, cudaStream_t st)
{
	switch (id) {
#define T(ID) case ID: return disk_shape<ID>(c, a, planes, ismax, two, st);
	DISK_IDS_MINE
#undef T
	}
	return -2;
}

#if MORSI_DISK_PART == 0
#ifdef MORSI_DISK_IDS   // single-unit development build: the other parts are empty
int morsi_disk_launch_part1(int, MorsiCtx *, const DiskArgs &, int, bool, bool, cudaStream_t) { return -2; }
int morsi_disk_launch_part2(int, MorsiCtx *, const DiskArgs &, int, bool, bool, cudaStream_t) { return -2; }
int morsi_disk_launch_part3(int, MorsiCtx *, const DiskArgs &, int, bool, bool, cudaStream_t) { return -2; }
#endif

template <int ID>
static bool shape_matches(const RowRunPlan &rr)
{
	if (rr.reach != Shape<ID>::R) return false;
	for (int i = 0; i <= 2 * rr.reach; i++)
		if (rr.hw[i] != Shape<ID>::hw(i)) return false;
	return true;
}

static int find_shape(const RowRunPlan &rr)
{
	if (!rr.ok) return -1;
#define T(ID) if (shape_matches<ID>(rr)) return ID;
	DISK_IDS
#undef T
	return -1;
}

static int launch_by_id(int id, MorsiCtx *c, const DiskArgs &a, int planes, bool ismax, bool two, cudaStream_t st)
{
	int rc = morsi_disk_launch_part0(id, c, a, planes, ismax, two, st);
	if (rc == -2) rc = morsi_disk_launch_part1(id, c, a, planes, ismax, two, st);
	if (rc == -2) rc = morsi_disk_launch_part2(id, c, a, planes, ismax, two, st);
	if (rc == -2) rc = morsi_disk_launch_part3(id, c, a, planes, ismax, two, st);
	if (rc == -2) return morsi_set_error(MORSI_ERR_INVALID, "no such shape %d", id);
	return rc;
}

// One reduction pass over `src` for output rows [row0,row0+rows) into `dst`.
static int disk_pass(MorsiCtx *c, int id, bool ismax, int epi, const MorsiJob &job, Band src, int src_rows,
		Band xop, Band other, float *dst, long long dst_pstride, int row0, int rows, int *flag)
{
	DiskArgs a;
	a.src = src; a.src_rows = src_rows; a.xop = xop; a.other = other;
	a.y = dst; a.y_pstride = dst_pstride; a.y_row0 = row0; a.y_rows = rows;
	a.w = job.w; a.h = job.h; a.epi = epi; a.flag = flag; a.band_rows = rows;
	a.pitch = job.pitch ? job.pitch : job.w;
	return launch_by_id(id, c, a, job.planes, ismax, false, job.stream);
}

// The passes of one operation over operands whose rows are P = job.pitch (or
// job.w) floats apart, 16-byte aligned.
static int run_disk_passes(MorsiCtx *c, int id, const DevElement *de, const MorsiJob &job, int *flag)
{
	const OpPlan plan = morsi_op_plan(job.op);
	const int P = job.pitch ? job.pitch : job.w;
	const int R = de->rowrun.reach;
	const Band none{nullptr, 0, 0};
	const Band xb{job.x, job.x_row0, job.x_pstride};
	int rc;

	// how many temporaries does the plan need?
	//   1 stage, one side      : 0      (erosion, dilation, i/egradient, i/eblur)
	//   1 stage, both sides    : 1      (gradient, laplacian, enhance, blur, cblur: a = erosion)
	//   2 stages               : 0      (opening, closing, tophat, bothat: fused)
	//   oscillation            : 3
	const bool both1 = plan.stages == 1 && plan.a_from && plan.b_from;
	const bool osc = plan.t_min && plan.t_max;
	if (plan.stages == 1 && !both1) {
		rc = disk_pass(c, id, plan.b_from != 0, plan.epi, job, xb, job.x_rows, xb, none,
				job.y, job.y_pstride, job.y_row0, job.y_rows, flag);
		return rc;
	}
	if (plan.stages == 2 && !osc) {
		// opening, closing, tophat, bothat: both stages in one kernel
		DiskArgs a;
		a.src = xb; a.src_rows = job.x_rows; a.other = none;
		a.xop = (plan.epi == EPI_X_SUB_B || plan.epi == EPI_A_SUB_X) ? xb : none;
		a.y = job.y; a.y_pstride = job.y_pstride; a.y_row0 = job.y_row0; a.y_rows = job.y_rows;
		a.w = job.w; a.h = job.h; a.epi = plan.epi; a.flag = flag; a.band_rows = job.y_rows;
		a.pitch = P;
		return launch_by_id(id, c, a, job.planes, plan.t_max != 0, true, job.stream);
	}
	if (both1) {
		// gradient, laplacian, enhance, blur, cblur: erosion and dilation in one pass (k_disk_both)
		static const bool two_pass = getenv("MORSI_DISK_BOTH") && !strcmp(getenv("MORSI_DISK_BOTH"), "0");   // A/B: the round-1 two-pass form
		// measured on B200 (profiles/r2_kdisk_both_swap_variants.txt): one pass wins for the small disks (disk7
		// gradient 0.247 -> 0.217 ms); for the 255-register shapes the two roles, coupled through the shared
		// input ring, lose to two passes (disk15 gradient on 40000x10000: 4.78 vs 3.68 ms)
		if (!two_pass && 4 * (2 * R + 2) <= 64) {
			DiskArgs a;
			a.src = xb; a.src_rows = job.x_rows; a.other = none; a.xop = xb;
			a.y = job.y; a.y_pstride = job.y_pstride; a.y_row0 = job.y_row0; a.y_rows = job.y_rows;
			a.w = job.w; a.h = job.h; a.epi = plan.epi; a.flag = flag; a.band_rows = job.y_rows;
			a.pitch = P;
			a.both = 1;
			return launch_by_id(id, c, a, job.planes, false, false, job.stream);
		}
	}
	// one temporary of the output band's size; chunk the band so that it stays below 512 MiB
	const long long budget = 512LL << 20;
	long long rows_fit = budget / ((long long)P * 4 * job.planes);
	if (rows_fit < 8 * R + 64) rows_fit = 8 * R + 64;
	const int chunk = (int)(rows_fit < job.y_rows ? rows_fit : job.y_rows);
	for (int r0 = 0; r0 < job.y_rows; r0 += chunk) {
		const int o0 = job.y_row0 + r0;
		const int orows = job.y_rows - r0 < chunk ? job.y_rows - r0 : chunk;
		float *ydst = job.y + (long long)r0 * P;
		void *p0; if ((rc = morsi_ws_get(c, job.lane, 0, (size_t)P * orows * job.planes * 4, &p0))) return rc;
		const long long tps = (long long)P * orows;
		if (both1) {
			rc = disk_pass(c, id, false, EPI_A, job, xb, job.x_rows, none, none, (float *)p0, tps, o0, orows, flag);
			if (rc) return rc;
			rc = disk_pass(c, id, true, plan.epi, job, xb, job.x_rows, xb, Band{(float *)p0, o0, tps},
					ydst, job.y_pstride, o0, orows, flag);
			if (rc) return rc;
			continue;
		}
		// oscillation = closing - opening (src/morsi.c:217-227): the fused closing into the
		// temporary, then the fused opening with the closing as the operand its epilogue
		// subtracts from (the tophat kernel, x := closing): two launches, one temporary
		DiskArgs a;
		a.src = xb; a.src_rows = job.x_rows; a.other = none; a.xop = none;
		a.y = (float *)p0; a.y_pstride = tps; a.y_row0 = o0; a.y_rows = orows;
		a.w = job.w; a.h = job.h; a.epi = EPI_A; a.flag = flag; a.band_rows = orows;
		a.pitch = P;
		if ((rc = launch_by_id(id, c, a, job.planes, true, true, job.stream))) return rc;       // closing
		a.xop = Band{(float *)p0, o0, tps};
		a.y = ydst; a.y_pstride = job.y_pstride; a.epi = EPI_X_SUB_B;
		if ((rc = launch_by_id(id, c, a, job.planes, false, true, job.stream))) return rc;      // closing - opening
	}
	return MORSI_OK;
}

// ---- widths / bases the TMA path cannot address directly -------------------------------
// Rows of a caller's plane are w floats apart; a tensor map needs a 16-byte
// pitch and base.  Such jobs run on pitched copies in the workspace: the copy
// in pads every row to a multiple of 4 floats with NaN (absent samples, like
// everything outside the image; the kernels still mask their temporaries with
// the true w), the copy out drops the pad.  16 B/sample of extra HBM traffic on
// kernels that are bound by the ALU pipe, instead of the exact kernels' n
// gathers per sample.
__global__ void __launch_bounds__(256) k_pitch_in(const float *x, long long x_pstride, float *d, long long d_pstride,
		int w, int pitch, int rows)
{
	const int plane = blockIdx.z;
	const int i = blockIdx.x * 256 + threadIdx.x;
	if (i >= pitch) return;
	for (int r = blockIdx.y; r < rows; r += gridDim.y)
		d[plane * d_pstride + (long long)r * pitch + i] = i < w ? __ldg(x + plane * x_pstride + (long long)r * w + i) : CUDART_NAN_F;
}
__global__ void __launch_bounds__(256) k_pitch_out(const float *s, long long s_pstride, float *y, long long y_pstride,
		int w, int pitch, int rows)
{
	const int plane = blockIdx.z;
	const int i = blockIdx.x * 256 + threadIdx.x;
	if (i >= w) return;
	for (int r = blockIdx.y; r < rows; r += gridDim.y)
		y[plane * y_pstride + (long long)r * w + i] = s[plane * s_pstride + (long long)r * pitch + i];
}

int morsi_run_disk(MorsiCtx *c, const DevElement *de, const MorsiJob &job, int *flag, int *handled)
{
	*handled = 0;
	const OpPlan plan = morsi_op_plan(job.op);
	if (plan.special) return MORSI_OK;
	const int id = find_shape(de->rowrun);
	if (id < 0) return MORSI_OK;
	const bool aligned = (job.w % 4 == 0) && (((uintptr_t)job.x) % 16 == 0) && (((uintptr_t)job.y) % 16 == 0)
		&& (job.x_pstride % 4 == 0) && (job.y_pstride % 4 == 0);
	int rc;
	if (aligned) {
		rc = run_disk_passes(c, id, de, job, flag);
		if (rc) return rc;
		*handled = 1;
		return MORSI_OK;
	}
	if (job.planes > 65535) return MORSI_OK;
	// pitched copies, in row chunks of at most ~256 MiB per copy
	const int R = de->rowrun.reach;
	const int halo = R * plan.stages;
	const int P = (job.w + 3) & ~3;
	long long rows_fit = (256LL << 20) / ((long long)P * 4 * job.planes) - 2 * halo;
	if (rows_fit < 8 * halo + 64) rows_fit = 8 * halo + 64;
	const int chunk = (int)(rows_fit < job.y_rows ? rows_fit : job.y_rows);
	for (int r0 = 0; r0 < job.y_rows; r0 += chunk) {
		const int o0 = job.y_row0 + r0;
		const int orows = job.y_rows - r0 < chunk ? job.y_rows - r0 : chunk;
		// input rows the chunk needs, clipped to what the caller's band holds
		int i0 = o0 - halo, i1 = o0 + orows + halo;
		if (i0 < job.x_row0) i0 = job.x_row0;
		if (i1 > job.x_row0 + job.x_rows) i1 = job.x_row0 + job.x_rows;
		const int irows = i1 - i0;
		void *px, *py;
		if ((rc = morsi_ws_get(c, job.lane, 6, (size_t)P * irows * job.planes * 4, &px))) return rc;
		if ((rc = morsi_ws_get(c, job.lane, 7, (size_t)P * orows * job.planes * 4, &py))) return rc;
		const dim3 gi((P + 255) / 256, irows < 1024 ? irows : 1024, job.planes);
		k_pitch_in<<<gi, 256, 0, job.stream>>>(job.x + (long long)(i0 - job.x_row0) * job.w, job.x_pstride,
				(float *)px, (long long)P * irows, job.w, P, irows);
		morsi_count_launch(1);
		MORSI_CU(cudaGetLastError());
		MorsiJob sub = job;
		sub.pitch = P;
		sub.x = (const float *)px; sub.x_row0 = i0; sub.x_rows = irows; sub.x_pstride = (long long)P * irows;
		sub.y = (float *)py; sub.y_row0 = o0; sub.y_rows = orows; sub.y_pstride = (long long)P * orows;
		if ((rc = run_disk_passes(c, id, de, sub, flag))) return rc;
		const dim3 go((job.w + 255) / 256, orows < 1024 ? orows : 1024, job.planes);
		k_pitch_out<<<go, 256, 0, job.stream>>>((const float *)py, (long long)P * orows,
				job.y + (long long)r0 * job.w, job.y_pstride, job.w, P, orows);
		morsi_count_launch(1);
		MORSI_CU(cudaGetLastError());
	}
	*handled = 1;
	return MORSI_OK;
}
#endif   // MORSI_DISK_PART == 0
