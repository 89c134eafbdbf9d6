/* A plain C11 caller of the C ABI (what a cgo / JNI / N-API binding compiles against): include/gpuntt_b200.h must be valid C, and the
 * library must link into a C program.  Built and run by tests/test_plan_host.py without a GPU: only calls that return before any
 * CUDA work are made (version, shapes, argument validation, empty batches). */
#include <stdio.h>
#include <string.h>

#include "gpuntt_b200.h"

int main(void)
{
    int n1 = 0, n2 = 0, bad = 0;
    gpuntt_b200_merge_desc d;
    gpuntt_b200_4step_desc f;
    printf("version %d\n", gpuntt_b200_version());
    bad += gpuntt_b200_version() != GPUNTT_B200_VERSION;
    bad += gpuntt_b200_4step_shape(24, &n1, &n2) != GPUNTT_B200_OK || n1 != 256 || n2 != 65536;
    bad += gpuntt_b200_4step_shape(11, &n1, &n2) != GPUNTT_B200_ERR_N_POWER;
    memset(&d, 0, sizeof d);
    d.element_bits = 64;
    d.direction = GPUNTT_B200_FORWARD;
    d.n_power = 16;
    d.ntt_layout = GPUNTT_B200_PER_POLYNOMIAL;
    d.reduction_poly = GPUNTT_B200_X_N_MINUS;
    d.modulus_value = 576460756061519873ull;
    d.batch_size = 0; /* an empty batch is a no-op, not an error */
    bad += gpuntt_b200_merge_ntt(&d) != GPUNTT_B200_OK;
    d.n_power = 29;
    bad += gpuntt_b200_merge_ntt(&d) != GPUNTT_B200_ERR_N_POWER;
    d.n_power = 16;
    d.ntt_layout = 7;
    bad += gpuntt_b200_merge_ntt(&d) != GPUNTT_B200_ERR_LAYOUT;
    d.ntt_layout = GPUNTT_B200_PER_POLYNOMIAL;
    d.batch_size = 4; /* null data pointers */
    bad += gpuntt_b200_merge_ntt(&d) != GPUNTT_B200_ERR_ARGUMENT;
    printf("last error: %s\n", gpuntt_b200_last_error());
    bad += gpuntt_b200_merge_ntt(NULL) != GPUNTT_B200_ERR_ARGUMENT;
    memset(&f, 0, sizeof f);
    f.element_bits = 64;
    f.n_power = 11;
    bad += gpuntt_b200_4step_ntt(&f) != GPUNTT_B200_ERR_N_POWER;
    bad += gpuntt_b200_transpose(16, NULL, NULL, 4, 4, 4, 1, NULL) != GPUNTT_B200_ERR_ARGUMENT;
    printf("%s\n", bad ? "FAILED" : "c caller ok");
    return bad;
}
