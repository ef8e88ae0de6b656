// Attention entry point of the C ABI: out = softmax(q k^T * scale + mask) v, head_dim 64.
//   HF modeling_clip.py:261-279 (eager_attention_forward), :318-331 (dispatch), :546-551 (causal + padding mask)
// Dispatches between the tcgen05 ping-pong kernel (attention_pp.cu; S <= 224, the vision tower), its two-launch
// key-range split (288 < S <= 384 without a mask, needs the caller's workspace) and the mma.sync kernel (attention.cu;
// short causal / masked text sequences, ViT-L/14's S = 257, and everything else).
#include <cstdio>
#include <cstdlib>

#include "../../include/vlmclip.h"
#include "common.cuh"

namespace vlmclip {
void count_launch(int n);
int attention_fwd_mma_sync(const void* qkv, void* out, const uint8_t* key_mask, int B, int S, int H, int causal,
                           float scale, cudaStream_t stream);
int attention_fwd_pingpong(const void* qkv, void* out, const uint8_t* key_mask, int B, int S, int H, int causal, float scale,
                           cudaStream_t stream);
int attention_fwd_pingpong_split(const void* qkv, void* out, float* workspace, int B, int S, int H, float scale,
                                 int variant, cudaStream_t stream);

}  // namespace vlmclip

using namespace vlmclip;

// VLMCLIP_ATTN_SPLIT: unset = auto (224 < S <= 288 on the mma.sync kernel, which measures 890 us against the
// split's 918 us at ViT-L/14, B=512; 288 < S <= 384 on split variant 3), 0 keeps every S > 224 on the mma.sync kernel,
// 1 / 2 / 3 / 4 force that variant of the key-range split (attention_pp.cu: attention_fwd_pingpong_split) for all
// 224 < S <= 384; A/B switch, read once
static int split_variant() {
  static const int variant = []() {
    const char* e = getenv("VLMCLIP_ATTN_SPLIT");
    return (e != nullptr && e[0] >= '0' && e[0] <= '4') ? e[0] - '0' : -1;
  }();
  return variant;
}

static bool split_eligible(int S, int causal, const uint8_t* key_mask) {
  const int v = split_variant();
  return v != 0 && S > (v < 0 ? 288 : 224) && S <= 384 && causal == 0 && key_mask == nullptr;
}

extern "C" int64_t vlmclip_attention_fwd_workspace(int B, int S, int H) {
  if (B <= 0 || S <= 0 || H <= 0 || !split_eligible(S, 0, nullptr)) return 0;
  return 2 * (int64_t)B * S * H;
}

extern "C" int vlmclip_attention_fwd_ws(const void* qkv, void* out, const uint8_t* key_mask, float* workspace, int B,
                                        int S, int H, int causal, float scale, void* stream) {
  VLMCLIP_CHECK_ARG(qkv && out, "attention: null pointer");
  VLMCLIP_CHECK_ARG(B > 0 && S > 0 && H > 0, "attention: bad dims B=%d S=%d H=%d", B, S, H);
  VLMCLIP_CHECK_ARG(S <= 512, "attention: S=%d exceeds the on-chip K/V limit (512)", S);
  VLMCLIP_CHECK_ARG((uintptr_t)qkv % 16 == 0 && (uintptr_t)out % 16 == 0, "attention: pointers must be 16-byte aligned");
  VLMCLIP_CHECK_ARG((uintptr_t)workspace % 8 == 0, "attention: workspace must be 8-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  // Dispatch.  S > 224 exceeds the two-S-buffer TMEM map of the tcgen05 kernel: up to S = 288 (ViT-L/14) the
  // mma.sync kernel is the faster of the two measured options, beyond that it runs as two key ranges merged in the second
  // launch's epilogue (needs the workspace).  For the short causal
  // text sequences (S = 77: 60 % of a 128-row tile would be padding, every chunk takes the masked path) the mma.sync
  // variant measures faster on B200 (43 us vs 98 us at B=256, H=8); VLMCLIP_ATTN_FORCE_TC=1 forces the tcgen05 kernel.
  static const bool force_tc = []() {
    const char* e = getenv("VLMCLIP_ATTN_FORCE_TC");
    return e != nullptr && e[0] == '1';
  }();
  static const bool force_mma = []() {  // A/B switch: every shape on the register-resident mma.sync kernel
    const char* e = getenv("VLMCLIP_ATTN_FORCE_MMA");
    return e != nullptr && e[0] == '1';
  }();
  if (force_mma) return attention_fwd_mma_sync(qkv, out, key_mask, B, S, H, causal, scale, s);
  if (workspace != nullptr && split_eligible(S, causal, key_mask))
    return attention_fwd_pingpong_split(qkv, out, workspace, B, S, H, scale, split_variant() < 0 ? 3 : split_variant(), s);
  if (S > 224 || (!force_tc && (causal != 0 || key_mask != nullptr) && S <= 128))
    return attention_fwd_mma_sync(qkv, out, key_mask, B, S, H, causal, scale, s);
  return attention_fwd_pingpong(qkv, out, key_mask, B, S, H, causal, scale, s);
}

extern "C" int vlmclip_attention_fwd(const void* qkv, void* out, const uint8_t* key_mask, int B, int S, int H,
                                     int causal, float scale, void* stream) {
  return vlmclip_attention_fwd_ws(qkv, out, key_mask, nullptr, B, S, H, causal, scale, stream);
}
