/* Minimal C host for libtapir_b200.so: what a non-Python caller (a C/C++ server, a JNI/cgo/FFI
 * shim) binds.  Only the C ABI of include/tapir_b200.h is used - no torch, no CUDA headers.
 *
 *   gcc -std=c99 -Iinclude examples/c_host.c -Ltapnet_b200 -ltapir_b200 \
 *       -Wl,-rpath,$PWD/tapnet_b200 -o c_host && ./c_host
 *
 * Without a GPU the size queries still work and the compute entry point reports a CUDA error
 * through the status code + tapir_last_error() (there is no CPU fallback to fall into).  With a
 * GPU the same call would need device pointers (cudaMalloc'ed by the host application). */
#include <stdio.h>
#include <stdint.h>

#include "tapir_b200.h"

int main(void) {
  printf("abi %d\n", tapir_abi_version());
  /* BASELINE config 2: 48 frames of 256x256, ExtraConvs on, 2 bf16 planes (bf16x3) */
  printf("backbone_ws %zu\n", tapir_backbone_workspace_bytes(48, 256, 256, 1, 2));
  printf("mixer_ws %zu\n", tapir_mixer_workspace_bytes(256 * 48, 2));
  printf("cost_volume_ws %zu\n", tapir_cost_volume_workspace_bytes(256, 48, 32, 32, 256));
  /* argument validation happens before any CUDA call */
  int rc = tapir_postprocess_occlusions(NULL, NULL, 0, NULL, NULL);
  printf("bad_args rc=%d msg=%s\n", rc, tapir_last_error());
  /* host pointers on purpose: this must fail with a status, never crash or compute on the CPU */
  static uint8_t frames[4 * 4 * 3];
  static float out[4 * 4 * 3];
  rc = tapir_ingest_frames(frames, 1, 4, 4, 0, 0, 4, 4, out, 4, 4, NULL);
  printf("host_pointers rc=%d msg=%s\n", rc, rc ? tapir_last_error() : "(ran on a GPU)");
  return 0;
}
