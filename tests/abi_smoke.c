/* Plain-C consumer of include/multinn_b200.h: proves the header is valid C99 and that a C program links against
 * libmultinn_sm100.so and reaches its argument checks without a GPU (tests/test_abi.py builds and runs this). */
#include <stdio.h>
#include <string.h>

#include "multinn_b200.h"

int main(void) {
  float buf[64];
  if (mnn_version() < 100) return 1;
  if (mnn_gemm_tc(NULL, 4, 0, NULL, 4, 0, NULL, 4, NULL, 1.0f, 0.0f, 8, 8, 8, 0, NULL) != MNN_ERR_ARG) return 2;
  if (strstr(mnn_last_error_string(), "null pointer") == NULL) return 3;
  if (mnn_rbm_gibbs_smem_bytes(84, 256) != (size_t)(2 * 84 * 256 + 8 * 4 * (84 + 256)) * sizeof(float)) return 4;
  if (mnn_rbm_gibbs_smem_bytes(420, 168) != 0) return 5;
  if (mnn_rbm_gibbs(buf, 84, buf, NULL, 0, NULL, 0, NULL, NULL, 0, 0ULL, 0ULL, buf, 84, buf, 84, NULL, 0, 8, 84, 256, 2,
                    NULL) != MNN_ERR_ARG) return 6;                 /* neither uniforms nor philox */
  if (mnn_lstm_tc_supported(8, 12) != 0 || mnn_lstm_tc_supported(8, 16) != 1) return 7;
  if (mnn_nade_logprob_fwd(NULL, buf, 4, 0, 0, buf, buf, buf, NULL, NULL, 0.0f, 8, 1, 84, 256, 0, NULL) != MNN_ERR_ARG)
    return 8;
  printf("abi smoke ok: version %d, launches so far %llu\n", mnn_version(), mnn_launch_count());
  return 0;
}
