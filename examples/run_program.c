/* A complete non-Python host for the hot path: loads a clip program saved by
 * vidsitu_b200.engine.ClipEngine.export_program (tools/export_program.py), feeds it uint8 frames and writes the
 * pooled features - the C counterpart of feat_extractor.py's inner loop (vidsitu_code/feat_extractor.py:86-112:
 * mdl.forward_encoder(batch) -> feature rows), through include/vidsitu_b200.h alone.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/run_program.c -o run_program \
 *       -L vidsitu_b200 -lvidsitu_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/vidsitu_b200
 *   ./run_program model.vsbprog frames.u8 feats.f32 [logits.f32]
 *
 * frames.u8: raw uint8 [n, T, H, W, 3] (the fast / single pathway window, dat_loader.py:474-476);
 * feats.f32: raw float32 [n, D]. */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "vidsitu_b200.h"

#define VSB(call)                                                           \
  do {                                                                      \
    if ((call) != VSB_OK) {                                                 \
      fprintf(stderr, "%s failed: %s\n", #call, vsb_last_error());          \
      return 1;                                                             \
    }                                                                       \
  } while (0)
#define CU(call)                                                            \
  do {                                                                      \
    cudaError_t e_ = (call);                                                \
    if (e_ != cudaSuccess) {                                                \
      fprintf(stderr, "%s failed: %s\n", #call, cudaGetErrorString(e_));    \
      return 1;                                                             \
    }                                                                       \
  } while (0)

static int dump(const char* path, const vsb_program* prog, const char* region) {
  void* dev;
  unsigned long long bytes;
  VSB(vsb_program_region(prog, region, &dev, &bytes));
  void* host = malloc(bytes);
  CU(cudaMemcpy(host, dev, bytes, cudaMemcpyDeviceToHost));
  FILE* f = fopen(path, "wb");
  if (!f || fwrite(host, 1, bytes, f) != bytes) {
    fprintf(stderr, "cannot write %s\n", path);
    return 1;
  }
  fclose(f);
  free(host);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: %s program.vsbprog frames.u8 feats.f32 [logits.f32]\n", argv[0]);
    return 2;
  }
  if (vsb_abi_version() != VSB_ABI_VERSION) {
    fprintf(stderr, "header is ABI %d, library is ABI %d\n", VSB_ABI_VERSION, vsb_abi_version());
    return 1;
  }
  /* workspace_bytes -> caller-owned device memory -> create */
  unsigned long long need = 0;
  VSB(vsb_program_file_device_bytes(argv[1], &need));
  void* mem = NULL;
  CU(cudaMalloc(&mem, need));
  vsb_program* prog = NULL;
  VSB(vsb_program_load(argv[1], mem, need, &prog));

  void* frames_dev;
  unsigned long long frames_bytes;
  VSB(vsb_program_region(prog, "frames", &frames_dev, &frames_bytes));
  void* frames = malloc(frames_bytes);
  FILE* f = fopen(argv[2], "rb");
  if (!f || fread(frames, 1, frames_bytes, f) != frames_bytes) {
    fprintf(stderr, "%s must hold %llu bytes of uint8 frames\n", argv[2], frames_bytes);
    return 1;
  }
  fclose(f);

  cudaStream_t stream;
  CU(cudaStreamCreate(&stream));
  VSB(vsb_program_capture(prog, stream)); /* optional: later runs replay one CUDA graph */
  CU(cudaMemcpyAsync(frames_dev, frames, frames_bytes, cudaMemcpyHostToDevice, stream));
  VSB(vsb_program_run(prog, stream)); /* forward */
  CU(cudaStreamSynchronize(stream));
  fprintf(stderr, "%d launches, %llu bytes of device memory\n", vsb_program_num_launches(prog), need);

  if (dump(argv[3], prog, "feats")) return 1;
  if (argc > 4 && dump(argv[4], prog, "logits")) return 1;
  vsb_program_destroy(prog);
  CU(cudaFree(mem));
  free(frames);
  return 0;
}
