// standalone probe of the TMA staging used by k_ccl_tile.  usage: tma_probe <variant>
//  0: 3-D tensor, box 48x33x1, negative start coords      1: same, start coords >= 0
//  2: 2-D tensor, box 48x33                                3: 3-D, box 64x33x1
//  5: 3-D, box 64x33x1, start x = x0-16 (16-byte aligned, may be negative)
//  4: libcu++ reference path (cuda::device::experimental::cp_async_bulk_tensor_2d_global_to_shared), box 48x33
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
namespace cde = cuda::device::experimental;
template <int BW, int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tmap, unsigned char *out, int x0, int y0, int fr) {
  __shared__ __align__(128) unsigned char t[33][BW];
  __shared__ __align__(8) unsigned long long mbar;
  const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar);
  const unsigned dst = (unsigned)__cvta_generic_to_shared(&t[0][0]);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((unsigned)(BW * 33)) : "memory");
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                   "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(x0), "r"(y0), "r"(fr), "r"(mb)
                   : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                   "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(x0), "r"(y0), "r"(mb)
                   : "memory");
  }
  asm volatile("{\n.reg .pred p;\nWAIT_TMA:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@!p bra WAIT_TMA;\n}\n" ::"r"(mb) : "memory");
  for (int i = threadIdx.x; i < 33 * BW; i += blockDim.x) out[i] = t[i / BW][i % BW];
}
__global__ void kref(const __grid_constant__ CUtensorMap tmap, unsigned char *out, int x0, int y0) {
  __shared__ __align__(128) unsigned char t[33][48];
  __shared__ cuda::barrier<cuda::thread_scope_block> bar;
  if (threadIdx.x == 0) {
    init(&bar, blockDim.x);
    cde::fence_proxy_async_shared_cta();
  }
  __syncthreads();
  cuda::barrier<cuda::thread_scope_block>::arrival_token tok;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&t, &tmap, x0, y0, bar);
    tok = cuda::device::barrier_arrive_tx(bar, 1, sizeof(t));
  } else {
    tok = bar.arrive();
  }
  bar.wait(std::move(tok));
  for (int i = threadIdx.x; i < 33 * 48; i += blockDim.x) out[i] = t[i / 48][i % 48];
}
int main(int argc, char **argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int Wp = 960, Hd = 540, B = 2;
  std::vector<unsigned char> h((size_t)Wp * Hd * B);
  for (size_t i = 0; i < h.size(); i++) h[i] = (unsigned char)(i * 7 + (i >> 8));
  unsigned char *d, *o;
  cudaMalloc(&d, h.size());
  cudaMalloc(&o, 33 * 64);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                               const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  const int rank = (variant == 2 || variant == 4) ? 2 : 3;
  const int bw = (variant == 3 || variant == 5) ? 64 : 48;
  CUtensorMap tm;
  const cuuint64_t dims[3] = {(cuuint64_t)Wp, (cuuint64_t)(rank == 2 ? Hd * B : Hd), (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)Wp, (cuuint64_t)Wp * Hd};
  const cuuint32_t box[3] = {(cuuint32_t)bw, 33, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult cr = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d encode: %d\n", variant, (int)cr);
  int bad = 0;
  const int tests[3][3] = {{variant == 1 ? 1 : 0, variant == 1 ? 1 : 0, 0}, {480, 256, 0}, {928, 500, 1}};
  for (auto &tc : tests) {
    const int cx = variant == 5 ? tc[0] - 16 : tc[0] - 1, cy = tc[1] - 1;
    if (variant == 4) kref<<<1, 256>>>(tm, o, cx, cy);
    else if (variant == 2) k<48, 2><<<1, 256>>>(tm, o, cx, cy, 0);
    else if (variant == 3 || variant == 5) k<64, 3><<<1, 256>>>(tm, o, cx, cy, tc[2]);
    else k<48, 3><<<1, 256>>>(tm, o, cx, cy, tc[2]);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  launch (%d,%d,%d): %s\n", cx, cy, tc[2], cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    unsigned char r[33 * 64];
    cudaMemcpy(r, o, 33 * bw, cudaMemcpyDeviceToHost);
    for (int yy = 0; yy < 33; yy++)
      for (int xx = 0; xx < bw; xx++) {
        int gx = cx + xx, gy = cy + yy;
        unsigned char want = 0;
        if (gx >= 0 && gx < Wp && gy >= 0 && gy < Hd * (rank == 2 ? B : 1)) want = h[(size_t)tc[2] * Wp * Hd + (size_t)gy * Wp + gx];
        if (r[yy * bw + xx] != want) bad++;
      }
  }
  printf("  mismatches: %d\n", bad);
  return bad != 0;
}
