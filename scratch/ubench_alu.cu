// ubench_alu.cu -- throughput of the min/max instructions the morsi kernels are
// bound by (lanes per clock per SM), alone and mixed with fma-pipe / LDS work.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/ubench_alu scratch/ubench_alu.cu
#include <cstdio>
#include <cuda_runtime.h>

#define REP 64
template <int MODE>
__global__ void __launch_bounds__(512) k(float *out, int iters, float s0, float s1, long long *clk)
{
	__shared__ float4 sm[1024];
	float a[8];
	for (int i = 0; i < 8; i++) a[i] = s0 * (threadIdx.x + i);
	float b = s1, c = s1 * 2.f, f = s0;
	int ia[8];
	for (int i = 0; i < 8; i++) ia[i] = threadIdx.x + i;
	int ib = (int)s1, ic = (int)s0;
	for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(s0, s1, s0, s1);
	__syncthreads();
	float4 acc4 = make_float4(0, 0, 0, 0);
	const float4 *sp = sm + (threadIdx.x & 511);
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int r = 0; r < REP / 8; r++) {
#pragma unroll
			for (int i = 0; i < 8; i++) {
				if (MODE == 0) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
				if (MODE == 1) asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
				if (MODE == 2) { asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
				                 asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(b), "f"(c)); }
				if (MODE == 3) { asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
				                 asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(ic) : "r"(ib), "r"(ib)); }
				if (MODE == 4) asm volatile("min.s32 %0, %0, %1;" : "+r"(ia[i]) : "r"(ib));
				if (MODE == 5) { asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
				                 if ((i & 3) == 0) { float4 t; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t.x),"=f"(t.y),"=f"(t.z),"=f"(t.w) : "r"((unsigned)__cvta_generic_to_shared(sp + 32 * ((i + r) & 15)))); acc4.x += t.x; } }
				if (MODE == 6) asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
				if (MODE == 7) asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(a[(i + 2) & 7]), "f"(a[(i + 5) & 7]));
				if (MODE == 8) { asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
				                 asm volatile("xor.b32 %0, %0, %1;" : "+r"(ic) : "r"(ia[i])); }
				if (MODE == 9) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
				if (MODE == 10) asm volatile("{.reg .s32 t; min.s32 t, %0, %1; min.s32 %0, t, %2;}" : "+r"(ia[i]) : "r"(ib), "r"(ic));
				if (MODE == 11) { asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
				                  asm volatile("{.reg .s32 t; min.s32 t, %0, %1; min.s32 %0, t, %2;}" : "+r"(ia[i]) : "r"(ib), "r"(ic)); }
				if (MODE == 12) { asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
				                  asm volatile("add.f32 %0, %0, %1;" : "+f"(f) : "f"(b)); }
				if (MODE == 13) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(ia[i]) : "r"(ib));
				if (MODE == 14) { asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
				                  asm volatile("min.f16x2 %0, %0, %1;" : "+r"(ia[i]) : "r"(ib)); }
				if (MODE == 15) { asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
				                  asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[(i+4)&7]) : "f"(c), "f"(b));
				                  asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(b), "f"(c));
				                  asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(ic) : "r"(ib), "r"(ib)); }
				if (MODE == 16) asm volatile("min.u32 %0, %0, %1;" : "+r"(ia[i]) : "r"(ib));
				if (MODE == 17) asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; selp.f32 %0, %0, %1, p;}" : "+f"(a[i]) : "f"(b));
			}
		}
	}
	long long t1 = clock64();
	float s = f + acc4.x;
	for (int i = 0; i < 8; i++) s += a[i] + ia[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s + ic;
	if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char *name, double ops_per_rep)
{
	float *out; long long *clk;
	cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&clk, 148 * 8);
	const int iters = 2000;
	k<MODE><<<148, 512>>>(out, 10, 1.f, 2.f, clk);
	k<MODE><<<148, 512>>>(out, iters, 1.f, 2.f, clk);
	long long h[148]; cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
	cudaError_t e = cudaDeviceSynchronize();
	double mx = 0; for (int i = 0; i < 148; i++) mx = h[i] > mx ? h[i] : mx;
	double lanes = 512.0 * iters * REP * ops_per_rep;
	printf("%-44s %8.1f lane-ops/clk/SM  (%s)\n", name, lanes / mx, cudaGetErrorString(e));
	cudaFree(out); cudaFree(clk);
}

int main()
{
	run<0>("FMNMX3 acc,b,c (3 regs)", 1);
	run<1>("FMNMX acc,b (2 regs)", 1);
	run<6>("FMNMX acc,acc' (dependent ring)", 1);
	run<7>("FMNMX3 3 distinct accs", 1);
	run<2>("FMNMX3 + FFMA 1:1 (count both)", 2);
	run<3>("FMNMX3 + IMAD 1:1 (count both)", 2);
	run<4>("IMNMX s32 2-input", 1);
	run<5>("FMNMX3 + LDS.128 4:1 (count FMNMX)", 1);
	run<8>("FMNMX3 + LOP3 1:1 (count both)", 2);
	run<9>("FADD", 1);
	run<10>("VIMNMX3 s32 3-input", 1);
	run<11>("FMNMX3 + VIMNMX3 1:1 (count both)", 2);
	run<12>("FMNMX3 + FADD 1:1 (count both)", 2);
	run<13>("HMNMX2 f16x2", 1);
	run<14>("FMNMX3 + HMNMX2 1:1 (count both)", 2);
	run<15>("2 FMNMX3 + FFMA + IMAD (count all 4)", 4);
	run<16>("IMNMX u32 2-input", 1);
	run<17>("FSETP+FSEL (count pair as 1)", 1);
	return 0;
}
