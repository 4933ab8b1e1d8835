/* A C caller of the boundary: includes the public header as C99, links against libvideoprism_b200.so and walks the
 * lifecycle a host-language binding (ctypes here; cgo / JNI / N-API elsewhere) would make.  Without a CUDA device every
 * compute entry point must fail loudly (VP_ERR_CUDA and a message, no CPU fallback); with one, a tiny encoder is created,
 * its expected parameter leaves are listed and the handle is destroyed.
 *
 *   gcc -std=c99 -Wall -Iinclude examples/abi_smoke.c -Lvideoprism-mlx_b200 -lvideoprism_b200 \
 *       -Wl,-rpath,$PWD/videoprism-mlx_b200 -o /tmp/abi_smoke && /tmp/abi_smoke
 */
#include <stdio.h>
#include <string.h>

#include "videoprism_b200.h"

int main(void) {
  vp_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.kind = VP_KIND_ENCODER;      /* the tiny encoder of the reference's encoders_test.py:115-181 */
  cfg.patch_size = 4;
  cfg.pos_emb_t = 4; cfg.pos_emb_h = 4; cfg.pos_emb_w = 4;
  cfg.model_dim = 32; cfg.num_spatial_layers = 2; cfg.num_temporal_layers = 2; cfg.num_heads = 2; cfg.mlp_dim = 64;
  cfg.atten_logit_cap = 50.0f;
  vp_handle* h = NULL;
  int rc = vp_create_ex(&cfg, -1, VP_FLAG_CHECK_FP32, &h);
  if (rc != VP_OK) {
    printf("vp_create_ex: status %d: %s\n", rc, vp_last_error(NULL));
    if (vp_device_sm_count() < 0 && rc == VP_ERR_CUDA) {
      printf("no CUDA device: failed loudly, as the contract says\n");
      return 0;
    }
    return 1;
  }
  printf("handle on device %d, flags %d, %d parameter leaves, first: %s\n", vp_handle_device(h), vp_handle_flags(h), vp_num_weights(h),
         vp_weight_key(h, 0));
  uint64_t ticket = 0;
  rc = vp_encoder_forward_host_async(h, &cfg, VP_F32, 1, 4, 16, 16, NULL, &cfg, NULL, VP_F32, NULL, &ticket);
  printf("forward before vp_finalize: status %d (%s)\n", rc, vp_last_error(h));   /* VP_ERR_INCOMPLETE */
  vp_destroy(h);
  return rc == VP_ERR_INCOMPLETE ? 0 : 1;
}
