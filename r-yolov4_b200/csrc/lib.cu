// Library-wide pieces of libryolo_b200.so: error string, version, device sanity.
#include "common.cuh"
#include <stdlib.h>
#include <string.h>

static thread_local char g_err[512] = "";

extern "C" {

void ryolo_set_error(const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

const char* ryolo_last_error(void) { return g_err; }

int ryolo_abi_version(void) { return 1; }

// Tuning / timing-experiment switches (process-wide).  Defaults come from the environment variable RYOLO_<KEY>
// (upper case) the first time a knob is read; ryolo_tune overrides them at run time.
//   halo      conv: 3x3/s1 taps read shifted views of one halo box (0 off | 1 | 2)
//   dbg       conv timing experiments, results are WRONG: 1 no stores, 2 no BN statistics, 4 no MMAs, 8 no A loads
//   wg_split  wgrad split-K shares of the tap groups: 1 = by measured cost per patch, 2 = by tap count, 0 = uniform
//   wg_dbg    wgrad timing experiments, results are WRONG: 1 no MMAs, 2 no X loads
//   wg_tapgrp wgrad: 1 = one MMA covers as many taps as fit N = 256, 0 = one tap per MMA
//   wg_trans  wgrad: 1 = layers with Cin <= 128 transpose their TMA boxes in smem and run K-major MMAs (correct, but
//             measured slower than the MN-major kernel: its raw ring is only 4 boxes deep), 0 (default) = MN-major
//   sw64      conv: 1 = Cin == 32 operands use 32-element (64-byte, SWIZZLE_64B) K blocks instead of overhanging 64-element boxes;
//             2 / 3 = also every layer with Cout <= 128 / every layer (half-size stages, twice the ring depth: measured
//             +9 % / +22 % on the conv stack: the per-stage hand-shake, not the ring depth, is what costs)
//   nacc      conv: 1 = as many TMEM accumulators in rotation as fit (up to 8), 0 = two
//   pdl       1 = the conv / wgrad / BatchNorm-backward / scale-shift-act kernels use programmatic dependent launch
//   bn_bwd    BatchNorm backward: 0 = original passes, 1 = low-register reduce (stores dY), 2 = reduce without the dY
//             store + an apply pass that recomputes the activation derivative, 3 (default) = variant 2's traffic with
//             register double-buffered rows, a resident-capacity grid and the 13-instruction Mish derivative
//   ssa       scale-shift-activation forward pass: 1 (default) = register double-buffered rows on a resident-capacity
//             grid, 0 = the four-loads-then-compute kernel
//   wg_boxes  wgrad: 16 KB shared-memory boxes per CTA (10..14).  14 (default) = the whole SM; 13 leaves 16 KB so that two
//             blocks of the HBM-bound BatchNorm-backward kernels fit on the same SM and run UNDER the side-stream wgrad
//   ew_regs   BatchNorm backward (variant 3): 0 (default) = the 128-register builds, 1 = the <=104-register builds (two
//             blocks + one wgrad CTA per SM) with the maximum shared-memory carveout, 2 = those with the default carveout
//             (co-residency experiment, with RYOLO_BWD_PRIO=1: measured neutral on the step)
//   nms_band  post_process: tiles (of 64 boxes) per band of the banded greedy NMS (pair work of the kept rows against
//             still-alive columns only; default 8, at most 16), 0 = the full N^2/2 mask followed by one scan
//   bn_fuse   BatchNorm backward (variant 3): 1 = layers of <= 24 M elements run both passes in ONE launch with a grid-wide
//             barrier in between (second pass reads L2), 0 (default) = two launches.  Measured: the barrier wait costs
//             what the second launch cost (53.3-53.5 vs 53.0-53.2 ms per step), see DESIGN.md section 8
//   kgrp      conv: K blocks per operand-ring stage (one barrier hand-shake and one commit per stage): 0 = one,
//             1 (default) = up to three for N <= 32 tiles (stages <= 40 KB, ring >= 3), 2 = any tile, <= 64 KB / ring
//             >= 2, 3 = <= 72 KB (see csrc/conv.cu launch(): only the narrow tiles gain)
//   pair      conv: 0 = off, n (default 256) = output-channel tiles of >= n columns with K >= 2304 run as clusters of two
//             CTAs that share every weight tile (each loads half of it, TMA multicast to both; csrc/conv.cu PAIR kernels);
//             + 1024 = also short K loops and small layers (parity tests)
//   wres      conv: largest weight matrix (KB, default 96) that is loaded once per CTA and stays resident in shared memory
//             (layers with one n-tile and >= 32 tiles per CTA; the ring then carries activation boxes only), 0 = off,
//             + 1024 = also layers with few tiles per CTA (parity tests)
//   wg_x32    wgrad: 1 (default) = the X boxes of Cin == 32 layers are 32 channels wide (SWIZZLE_64B MN-major atoms), 0 = 64-channel
//             boxes overhanging the tensor
//   epi_tma   conv bf16 epilogue: 0 per-thread 16-byte stores | 1 TMA slab stores | 2 (default) + TMA reduce-add for
//             dgrad's accumulation;  epi_maxbn: widest tile that always takes the slab path (wider ones only with K <= 1152)
static const char* const kKnobNames[RYOLO_KNOB_COUNT] = {"halo", "dbg", "wg_split", "wg_dbg", "epi_tma", "epi_maxbn",
                                                          "wg_tapgrp", "bn_bwd", "wg_trans", "sw64", "nacc", "pdl", "ssa", "wg_boxes", "ew_regs", "nms_band", "bn_fuse", "kgrp", "pair", "wres", "wg_x32"};
static const int kKnobDefaults[RYOLO_KNOB_COUNT] = {0, 0, 1, 0, 2, 128, 1, 3, 0, 1, 1, 1, 1, 14, 0, 8, 0, 1, 256, 96, 1};
static int g_knobs[RYOLO_KNOB_COUNT];
static bool g_knob_set[RYOLO_KNOB_COUNT];

int ryolo_knob(int id) {
  if (id < 0 || id >= RYOLO_KNOB_COUNT) return 0;
  if (!g_knob_set[id]) {
    char env[64] = "RYOLO_";
    size_t n = strlen(env);
    for (const char* c = kKnobNames[id]; *c && n + 1 < sizeof(env); c++) env[n++] = (char)((*c >= 'a' && *c <= 'z') ? *c - 32 : *c);
    env[n] = 0;
    const char* e = getenv(env);
    g_knobs[id] = e ? atoi(e) : kKnobDefaults[id];
    g_knob_set[id] = true;
  }
  return g_knobs[id];
}

int ry_sm_count(void) {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int ryolo_tune(const char* key, int value) {
  for (int i = 0; i < RYOLO_KNOB_COUNT; i++) {
    if (key && strcmp(key, kKnobNames[i]) == 0) { g_knobs[i] = value; g_knob_set[i] = true; return RYOLO_OK; }
  }
  ryolo_set_error("ryolo_tune: unknown key");
  return RYOLO_ERR_INVALID;
}

// Returns RYOLO_OK only on a compute-capability 10.x device (the library ships sm_100a SASS only).
int ryolo_check_device(int device) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, device);
  if (e != cudaSuccess) { ryolo_set_error(cudaGetErrorString(e)); return RYOLO_ERR_CUDA; }
  if (p.major != 10) { ryolo_set_error("libryolo_b200 is built for sm_100a (B200) only"); return RYOLO_ERR_INVALID; }
  return RYOLO_OK;
}

}  // extern "C"
