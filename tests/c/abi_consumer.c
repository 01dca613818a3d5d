/* A plain C consumer of include/gpud_b200.h, linked against libgpud_b200.so the way a cgo file would be (tests/test_abi_cpu.py builds
 * and runs it).  Host-only entry points answer without a GPU; compute entry points fail loudly instead of falling back to the CPU. */
#include <stdio.h>
#include <string.h>

#include "gpud_b200.h"

int main(void) {
  char buf[512];
  int fails = 0;
  if (gpud_abi_version() <= 0) { printf("abi version\n"); ++fails; }
  if (gpud_sizeof(0) != (int32_t)sizeof(gpud_xid_hit) || gpud_sizeof(1) != (int32_t)sizeof(gpud_fabric_raw) ||
      gpud_sizeof(10) != (int32_t)sizeof(gpud_temperature) || gpud_sizeof(12) != (int32_t)sizeof(gpud_event_row)) { printf("layout\n"); ++fails; }
  /* xid/health_state_test.go:452-466 */
  if (gpud_xid_build_message(149, 37, 0, "", "PCI:0000:00:00", NULL, buf, (int32_t)sizeof buf) < 0 ||
      strcmp(buf, "XID 149.37 (err status 0x00000000) NVLINK_NETIR_ERROR detected on GPU PCI:0000:00:00") != 0) { printf("message: %s\n", buf); ++fails; }
  if (gpud_product_mem_caps("NVIDIA B200") != 7 || gpud_product_fm_supported("NVIDIA H100 PCIe") != 0) { printf("product\n"); ++fails; }
  {
    gpud_temperature t;
    int32_t bits = -1;
    memset(&t, 0, sizeof t);
    t.current_gpu_core_c = 90; t.threshold_gpu_max_c = 88;
    if (gpud_temperature_check(&t, 0, &bits) != GPUD_OK || bits != 1) { printf("temperature\n"); ++fails; }
  }
  {
    gpud_ctx* ctx = NULL;
    int32_t dev = 0;
    const int32_t rc = gpud_ctx_create(&dev, 1, &ctx);
    if (rc == GPUD_OK) { printf("gpu present: ctx ok\n"); gpud_ctx_destroy(ctx); }
    else if (rc != GPUD_E_CUDA) { printf("ctx_create rc %d\n", rc); ++fails; }
    else printf("no gpu: ctx_create -> GPUD_E_CUDA\n");
  }
  printf(fails ? "FAIL %d\n" : "OK\n", fails);
  return fails;
}
