extern "C" int xt_predict(xt_ctx* ctx, const xt_params* p, double* const* out) {
  (void)p; (void)out;
  set_error(ctx, "xt_predict: not implemented yet");
  return XT_ERR_STATE;
}
