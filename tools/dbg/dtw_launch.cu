#include "../../abnet3_b200/csrc/abn_align.cu"
namespace abn { char *err_buf(){ static char b[512]; return b;} int set_error(int c, const char*f, ...){ va_list ap; va_start(ap,f); vfprintf(stderr,f,ap); va_end(ap); fprintf(stderr,"\n"); return c;} int require_sm100(){return 0;} }
int main(){
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, abn::dtw_skew_kernel<2>);
  printf("attr: %s regs %d local %zu shared %zu maxdyn %d maxthreads %d ptx %d bin %d\n", cudaGetErrorString(e), fa.numRegs, fa.localSizeBytes, fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes, fa.maxThreadsPerBlock, fa.ptxVersion, fa.binaryVersion);
  abn::AlignArgs a{}; int *d; cudaMalloc(&d, 4096); cudaMemset(d,0,4096); a.class_off = d; a.order=d; a.w0=0; a.w1=0;
  printf("setattr: %s\n", cudaGetErrorString(cudaFuncSetAttribute(abn::dtw_skew_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 11520)));
  int per_sm=0; printf("occ: %s ", cudaGetErrorString(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, abn::dtw_skew_kernel<2>, 128, 11520))); printf("%d\n", per_sm);
  cudaStream_t st; cudaStreamCreate(&st);
  abn::dtw_skew_kernel<2><<<100,128,11520,st>>>(a,0,1,40);
  printf("launch: %s\n", cudaGetErrorString(cudaGetLastError()));
  abn::dtw_skew_kernel<1><<<100,128,11520>>>(a,0,1,40);
  printf("launch G1: %s\n", cudaGetErrorString(cudaGetLastError()));
  printf("sync: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
