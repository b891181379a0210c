// extern "C" entry points for the low-level building blocks (packing, tensor-core GEMM, SIMT GEMM).
#include "../../include/vame_b200.h"
#include "api_common.h"
#include "common.cuh"
#include "kernels.h"

namespace vb {
thread_local char g_err[512] = {0};
long g_launch_count = 0;
int g_opt_pdl = 1;
int g_opt_streams = 1;
int g_opt_warps16 = 1;
int g_opt_slice16 = 1;
int g_opt_m64 = 1;
int g_opt_flags = 0;           // measured: no gain (the gpu-scope publish costs what the kernel-completion flush costs)
int g_opt_persistent = 2;      // bit 0: forward sweeps, bit 1: backward sweeps as persistent cluster kernels
int g_opt_rw = 3;              // bit 0 / bit 1: forward / backward sweeps by the resident-weight cluster kernels when applicable
int g_opt_rw_waves = 2;      // measured: two waves of resident-weight clusters beat the slice kernels at B = 512 (C5 +31 %, C3 +22 %)
int g_opt_rw2 = 1;
int g_opt_rows = 4;             // row-resident forward sweep (gru_rows.cu) for large inference batches: 0 off, 1 = 8 gate-math warps, 4 = 16 (measured: 3.03 vs 3.08 M windows/s at C4)
int g_opt_side_sms = 0;
int g_opt_side_split = 2;
int g_opt_rw_ng = 0;            // 16-row groups per cluster of the H = 256 rw kernels: 0 = automatic, 1 or 2 forced
int g_opt_rw_exp = 0;
int g_opt_rw_priv = 1;
int g_opt_rw_sw = 7;             // store warps: bit 0 forward sweeps, bit 1 BPTT sweeps with 16 rows per cluster (round 2, after the per-role loops: 3.68 -> 3.22 us per step, C2 +3.4 %), bit 2 BPTT sweeps with 2 x 16 rows
unsigned long long* g_dbg_buffer = nullptr;
}

extern "C" {

const char* vame_last_error(void) { return vb::g_err; }

int vame_abi_version(void) { return VAME_B200_ABI_VERSION; }

long vame_launch_count(void) { return vb::g_launch_count; }

int vame_set_debug_buffer(void* device_u64x16) {
  vb::g_dbg_buffer = (unsigned long long*)device_u64x16;
  return 0;
}

int vame_get_option(const char* name) {
  if (!name) return -1;
  if (strcmp(name, "pdl") == 0) return vb::g_opt_pdl;
  if (strcmp(name, "persistent") == 0) return vb::g_opt_persistent;
  if (strcmp(name, "flags") == 0) return vb::g_opt_flags;
  if (strcmp(name, "streams") == 0) return vb::g_opt_streams;
  if (strcmp(name, "warps16") == 0) return vb::g_opt_warps16;
  if (strcmp(name, "slice16") == 0) return vb::g_opt_slice16;
  if (strcmp(name, "m64") == 0) return vb::g_opt_m64;
  if (strcmp(name, "rw") == 0) return vb::g_opt_rw;
  if (strcmp(name, "rw_waves") == 0) return vb::g_opt_rw_waves;
  if (strcmp(name, "rw2") == 0) return vb::g_opt_rw2;
  if (strcmp(name, "rw_ng") == 0) return vb::g_opt_rw_ng;
  if (strcmp(name, "rows") == 0) return vb::g_opt_rows;
  if (strcmp(name, "side_sms") == 0) return vb::g_opt_side_sms;
  if (strcmp(name, "side_split") == 0) return vb::g_opt_side_split;
  if (strcmp(name, "rw_priv") == 0) return vb::g_opt_rw_priv;
  if (strcmp(name, "rw_sw") == 0) return vb::g_opt_rw_sw;
  if (strcmp(name, "rw_timeouts") == 0) return (int)vb::rw_timeouts();
  if (strcmp(name, "rows_timeouts") == 0) return (int)vb::rows_timeouts();
  if (strncmp(name, "rw_timeout_info", 15) == 0) return vb::rw_timeout_info(name[15] ? name[15] - '0' : 0);
  return -1;
}

int vame_set_option(const char* name, int value) {
  VB_REQUIRE(name, "vame_set_option: null name");
  if (strcmp(name, "pdl") == 0) {
    vb::g_opt_pdl = value ? 1 : 0;
    return 0;
  }
  if (strcmp(name, "persistent") == 0) {
    vb::g_opt_persistent = value;
    return 0;
  }
  if (strcmp(name, "flags") == 0) {
    vb::g_opt_flags = value ? 1 : 0;
    return 0;
  }
  if (strcmp(name, "warps16") == 0) {
    vb::g_opt_warps16 = value ? 1 : 0;
    return 0;
  }
  if (strcmp(name, "streams") == 0) {
    vb::g_opt_streams = value ? 1 : 0;
    return 0;
  }
  if (strcmp(name, "slice16") == 0) {
    vb::g_opt_slice16 = value ? 1 : 0;
    return 0;
  }
  if (strcmp(name, "m64") == 0) {
    vb::g_opt_m64 = value ? 1 : 0;
    return 0;
  }
  if (strcmp(name, "rw") == 0) {
    vb::g_opt_rw = value & 3;
    return 0;
  }
  if (strcmp(name, "rw_timeouts_reset") == 0) {
    vb::rw_timeouts_reset();
    return 0;
  }
  if (strcmp(name, "rw_priv") == 0) {
    vb::g_opt_rw_priv = value ? 1 : 0;
    return 0;
  }
  if (strcmp(name, "rw_sw") == 0) {
    vb::g_opt_rw_sw = value;
    return 0;
  }
  if (strcmp(name, "rw_exp") == 0) {
    vb::g_opt_rw_exp = value;
    return 0;
  }
  if (strcmp(name, "rw2") == 0) {
    vb::g_opt_rw2 = value ? 1 : 0;
    return 0;
  }
  if (strcmp(name, "rw_waves") == 0) {
    vb::g_opt_rw_waves = value < 1 ? 1 : value;
    return 0;
  }
  if (strcmp(name, "rw_ng") == 0) {
    vb::g_opt_rw_ng = (value == 1 || value == 2) ? value : 0;
    return 0;
  }
  if (strcmp(name, "side_sms") == 0) {
    vb::g_opt_side_sms = value < 0 ? 0 : value;
    return 0;
  }
  if (strcmp(name, "side_split") == 0) {
    vb::g_opt_side_split = value == 1 ? 1 : 2;
    return 0;
  }
  if (strcmp(name, "rows") == 0) {
    vb::g_opt_rows = value;            // 0 off, 1-3: 8 gate-math warps, >= 4: 16 gate-math warps
    return 0;
  }
  return vb::fail("vame_set_option: unknown option");
}

size_t vame_p16_bytes(int rows, int k, int row_block) { return vb::p16_bytes(rows, k, row_block); }

int vame_pack_p16(const float* src, long ld, int transposed, int rows, int k, int rows_src, int k_src, const int* row_map,
                  const int* col_map, int row_block, void* out, void* stream) {
  VB_REQUIRE(src && out, "vame_pack_p16: null pointer");
  VB_REQUIRE(row_block > 0 && row_block % 8 == 0, "vame_pack_p16: row_block must be a positive multiple of 8");
  VB_REQUIRE(rows > 0 && k > 0, "vame_pack_p16: empty matrix");
  vb::launch_pack_p16(src, ld, transposed, rows, k, rows_src, k_src, row_map, col_map, row_block, out, (cudaStream_t)stream);
  return vb::check_launch("vame_pack_p16");
}

int vame_gemm_p16(const void* a_p, int a_nkc, const void* b_p, int b_nkc, int M, int N, float* C, long ldc, const float* bias,
                  int accumulate, int splits, void* stream) {
  VB_REQUIRE(a_p && b_p && C, "vame_gemm_p16: null pointer");
  VB_REQUIRE(a_nkc == b_nkc && a_nkc > 0, "vame_gemm_p16: A and B must have the same number of K chunks");
  VB_REQUIRE(M > 0 && N > 0, "vame_gemm_p16: empty output");
  VB_REQUIRE(splits >= 1 && (splits == 1 || accumulate), "vame_gemm_p16: split-K requires accumulate=1");
  vb::GemmArgs g{};
  g.a[0] = {a_p, a_nkc, a_nkc};
  g.a[1] = {nullptr, 0, 0};
  g.b[0] = {b_p, b_nkc, b_nkc};
  g.b[1] = {nullptr, 0, 0};
  g.M = M; g.N = N; g.C = C; g.ldc = ldc; g.bias = bias; g.atomic = accumulate; g.splits = splits;
  vb::launch_gemm_p16(g, (cudaStream_t)stream);
  return vb::check_launch("vame_gemm_p16");
}

}  // extern "C"
