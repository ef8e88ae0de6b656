// Native launch sequence of one frozen CLIP tower (HF modeling_clip.py:355-386 CLIPEncoderLayer x L), LayerNorm folded
// into the following GEMM.  The kernels are the ones the per-op entry points launch; this entry point only removes
// the host-side cost of issuing them one by one from Python (about 50 us per op, 170 ops per step: the interpreter
// was slower than the GPU), so the frozen towers are two calls per step.
#include "../../include/vlmclip.h"
#include "common.cuh"

extern "C" int vlmclip_encoder_fwd(const vlmclip_layer_t* layers, int n_layers, void* x, void* x_lo, void* qkv, void* att, void* hid,
                                   float* stats, float* part, int32_t* row_counters, const uint8_t* key_mask, int B, int S,
                                   int H, int D, int F, float eps, int causal, int act, void* stream) {
  using namespace vlmclip;
  VLMCLIP_CHECK_ARG(layers && n_layers > 0 && x && qkv && att && hid && stats && part, "encoder_fwd: null pointer");
  VLMCLIP_CHECK_ARG(B > 0 && S > 0 && H > 0 && D == H * 64 && F > 0 && D % 32 == 0, "encoder_fwd: bad dims");
  const int M = B * S;
  const int npart = D / 32;
  // two-term residual stream: x = hi plane, x_lo = lo plane (any distance apart, as long as it is a multiple of 8 elements)
  const int64_t plane = x_lo != nullptr ? (static_cast<const __nv_bfloat16*>(x_lo) - static_cast<const __nv_bfloat16*>(x)) : 0;
  VLMCLIP_CHECK_ARG(x_lo == nullptr || (plane >= (int64_t)M * D && plane % 8 == 0),
                    "encoder_fwd: x_lo must follow x by at least B*S*D elements (multiple of 8)");
  const float scale = 0.125f;  // head_dim 64
  int rc = 0;
  // with row_counters the residual GEMMs (out-proj, fc2) leave finished (mean, rstd) rows in `stats` themselves
  float* fused_stats = row_counters != nullptr ? stats : nullptr;
  for (int l = 0; l < n_layers; ++l) {
    const vlmclip_layer_t& L = layers[l];
    // LN1 statistics of the residual stream: explicit pass for the first layer (the embedding kernels emit no
    // partials), otherwise the combine of the per-32-column partials the previous fc2 epilogue left
    rc = l == 0 ? vlmclip_row_stats_bf16(x, D, stats, M, D, eps, stream)
                : (fused_stats != nullptr ? 0 : vlmclip_ln_partials_to_stats(part, stats, M, npart, eps, stream));
    if (rc) return rc;
    rc = vlmclip_gemm_bf16(x, D, L.qkv_w, D, qkv, 3 * (int64_t)D, L.qkv_b, nullptr, 0, stats, L.qkv_c, nullptr, 0, eps,
                           nullptr, nullptr, nullptr, M, 3 * D, D, VLMCLIP_ACT_NONE, 0, stream);
    if (rc) return rc;
    // `part` is dead between the LN1 combine above and the out-proj epilogue below and holds M * D / 16 >= 2 * M * H
    // floats: it doubles as the scratch of the key-range split (S = 257)
    rc = vlmclip_attention_fwd_ws(qkv, att, key_mask, part, B, S, H, causal, scale, stream);
    if (rc) return rc;
    // x += out_proj(att), in place; the epilogue leaves the LN2 partials (and, fused, the LN2 statistics)
    rc = x_lo != nullptr ? vlmclip_gemm_bf16_res2(att, D, L.out_w, D, x, D, plane, L.out_b, part, fused_stats, row_counters,
                                                  eps, M, D, D, stream)
                         : vlmclip_gemm_bf16(att, D, L.out_w, D, x, D, L.out_b, x, D, nullptr, nullptr, nullptr, 0, eps, part,
                                             fused_stats, row_counters, M, D, D, VLMCLIP_ACT_NONE, 0, stream);
    if (rc) return rc;
    if (fused_stats == nullptr) {
      rc = vlmclip_ln_partials_to_stats(part, stats, M, npart, eps, stream);
      if (rc) return rc;
    }
    rc = vlmclip_gemm_bf16(x, D, L.fc1_w, D, hid, F, L.fc1_b, nullptr, 0, stats, L.fc1_c, nullptr, 0, eps, nullptr, nullptr,
                           nullptr, M, F, D, act, 0, stream);
    if (rc) return rc;
    rc = x_lo != nullptr ? vlmclip_gemm_bf16_res2(hid, F, L.fc2_w, F, x, D, plane, L.fc2_b, part, fused_stats, row_counters,
                                                  eps, M, D, F, stream)
                         : vlmclip_gemm_bf16(hid, F, L.fc2_w, F, x, D, L.fc2_b, x, D, nullptr, nullptr, nullptr, 0, eps, part,
                                             fused_stats, row_counters, M, D, F, VLMCLIP_ACT_NONE, 0, stream);
    if (rc) return rc;
  }
  return 0;
}
