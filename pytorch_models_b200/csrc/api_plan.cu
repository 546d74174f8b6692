// b200enc_run_ops: replay of a recorded launch sequence (include/b200enc.h, "Launch plan"). Pure host code: every
// element is handed to the entry point it names, so validation, tensor-map caching and error reporting are the ones
// of the individual calls.
#include "../../include/b200enc.h"
#include "host_util.h"

namespace {

int run_one(const b200enc_op& o, void* s) {
  const void* const* p = o.p;
  const long long* i = o.i;
  switch (o.kind) {
    case B200ENC_OP_LINEAR:
      return b200enc_linear(&o.linear, s);
    case B200ENC_OP_PATCH_EMBED16:
      return b200enc_patch_embed16(&o.linear, int(i[0]), int(i[1]), s);
    case B200ENC_OP_ATTENTION:
      return b200enc_attention(p[0], i[0], int(i[1]), p[1], p[2], i[2], int(i[3]), const_cast<void*>(p[3]), i[4],
                               int(i[5]), int(i[6]), int(i[7]), int(i[8]), int(i[9]), int(i[10]), o.f[0], int(i[11]), s);
    case B200ENC_OP_ATTENTION_BIAS:
      return b200enc_attention_bias(p[0], i[0], int(i[1]), p[1], p[2], i[2], int(i[3]), const_cast<void*>(p[3]), i[4],
                                    int(i[5]), int(i[6]), int(i[7]), int(i[8]), int(i[9]), int(i[10]), o.f[0],
                                    int(i[11]), static_cast<const float*>(p[4]), i[12], i[13], i[14], s);
    case B200ENC_OP_LAYERNORM:
      return b200enc_layernorm(p[0], i[0], static_cast<const float*>(p[1]), static_cast<const float*>(p[2]), o.f[0],
                               int(i[1]), int(i[2]), const_cast<void*>(p[3]), i[3],
                               static_cast<float*>(const_cast<void*>(p[4])), s);
    case B200ENC_OP_ROW_STATS:
      return b200enc_row_stats(p[0], i[0], o.f[0], int(i[1]), int(i[2]), static_cast<float*>(const_cast<void*>(p[1])), s);
    case B200ENC_OP_MEAN_TOKENS:
      return b200enc_mean_tokens(p[0], i[0], i[1], int(i[2]), int(i[3]), int(i[4]), const_cast<void*>(p[1]), i[5], s);
    case B200ENC_OP_PATCH_ROWS:
      return b200enc_patch_rows(p[0], int(i[0]), int(i[1]), int(i[2]), int(i[3]), int(i[4]), int(i[5]),
                                const_cast<void*>(p[1]), s);
    case B200ENC_OP_CLS_ROWS:
      return b200enc_cls_rows(p[0], int(i[0]), int(i[1]), const_cast<void*>(p[1]), i[2], s);
    case B200ENC_OP_EMBED_ROWS:
      return b200enc_embed_rows(static_cast<const long long*>(p[0]), i[0], int(i[1]), p[1], p[2], int(i[2]), int(i[3]),
                                int(i[4]), const_cast<void*>(p[3]), s);
    case B200ENC_OP_TIME_ROWS:
      return b200enc_time_rows(p[0], int(i[0]), int(i[1]), int(i[2]), int(i[3]), const_cast<void*>(p[1]), s);
    default:
      return b200::set_error(-1, "b200enc_run_ops: unknown op kind %d", o.kind);
  }
}

}  // namespace

extern "C" int b200enc_run_ops(const b200enc_op* ops, int n_ops, int* failed_op, void* stream) {
  if (failed_op) *failed_op = -1;
  B200_CHECK_ARG(n_ops >= 0 && (ops != nullptr || n_ops == 0), "b200enc_run_ops: null op array");
  for (int k = 0; k < n_ops; ++k) {
    if (int rc = run_one(ops[k], stream)) {
      if (failed_op) *failed_op = k;
      return rc;
    }
  }
  return 0;
}
