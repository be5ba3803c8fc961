// N-API addon over the C ABI of include/hzwitness.h: what `witness_calculator.js` (the circom_runtime
// `WitnessCalculator` surface the reference's tests and tools drive: /root/reference/test/helpers/helpers.js:142-155,
// test/rollup-main.test.js:52,873, tools/helpers/actions.js:139-143) calls into.  Every line of circuit logic lives
// behind the C ABI; this file only moves buffers and turns error codes into the runtime's exceptions.
//
//   create(kind, nTx, nLevels, maxL1Tx, maxFeeTx, device, nGpus, symPath|null) -> external(ctx)
//   nvars(ctx), ninputs(ctx)                          -> number
//   inputIndex(ctx, name)                             -> [offset, length]     (throws "Signal <name> is not an input of the circuit.")
//   calculate(ctx, inputs: Uint8Array, sanity: bool)  -> Promise<Uint8Array>  (hzw_calculate on a libuv worker thread)
//   getSignal(ctx, witness: Uint8Array, name)         -> Uint8Array(32)
//   signalIndex(ctx, name)                            -> number
//   writeWtns(ctx, witness: Uint8Array, path)         -> undefined
//   checkConstraints(ctx, r1csPath, witness|null)     -> [constraints, violated, firstViolated]
//
// Build: `npm install` in this directory (binding.gyp links ../../circuits_b200/libhzwitness.so).
#include <node_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include "../../include/hzwitness.h"

#define NAPI_OK(call)                                                                  \
  do {                                                                                 \
    if ((call) != napi_ok) {                                                           \
      napi_throw_error(env, nullptr, "hzwitness: N-API call failed: " #call);          \
      return nullptr;                                                                  \
    }                                                                                  \
  } while (0)

static std::string hz_decimal(const uint8_t le[32]) {
  uint32_t t[8];
  for (int i = 0; i < 8; i++) t[i] = (uint32_t)le[4 * i] | ((uint32_t)le[4 * i + 1] << 8) | ((uint32_t)le[4 * i + 2] << 16) | ((uint32_t)le[4 * i + 3] << 24);
  std::string s;
  bool nz = true;
  while (nz) {
    uint64_t rem = 0;
    nz = false;
    for (int i = 7; i >= 0; i--) {
      uint64_t cur = (rem << 32) | t[i];
      t[i] = (uint32_t)(cur / 10);
      rem = cur % 10;
      if (t[i]) nz = true;
    }
    s.insert(s.begin(), (char)('0' + rem));
  }
  return s;
}
// the message circom_runtime throws for a violated `===` (tests match its prefix: test/rollup-main.test.js:683,876)
static std::string hz_error_message(int st, const hzw_error& e) {
  if (st == HZW_E_CONSTRAINT) return "Constraint doesn't match " + hz_decimal(e.lhs) + " != " + hz_decimal(e.rhs) + " " + e.where;
  if (st == HZW_E_DEVICE) return "hzwitness: CUDA device error (this library has no CPU path)";
  char b[64];
  snprintf(b, sizeof b, "hzwitness: error %d", st);
  return b;
}

static hzw_ctx* hz_ctx_arg(napi_env env, napi_value v) {
  void* p = nullptr;
  if (napi_get_value_external(env, v, &p) != napi_ok || !p) { napi_throw_type_error(env, nullptr, "hzwitness: not a context"); return nullptr; }
  return static_cast<hzw_ctx*>(p);
}
static bool hz_string_arg(napi_env env, napi_value v, std::string& out) {
  size_t n = 0;
  if (napi_get_value_string_utf8(env, v, nullptr, 0, &n) != napi_ok) { napi_throw_type_error(env, nullptr, "hzwitness: string expected"); return false; }
  out.resize(n);
  return napi_get_value_string_utf8(env, v, &out[0], n + 1, &n) == napi_ok;
}
// pointer and length of a Uint8Array / Buffer argument
static bool hz_bytes_arg(napi_env env, napi_value v, uint8_t** data, size_t* len) {
  bool is_ta = false;
  if (napi_is_typedarray(env, v, &is_ta) != napi_ok || !is_ta) { napi_throw_type_error(env, nullptr, "hzwitness: Uint8Array expected"); return false; }
  napi_typedarray_type type; napi_value ab; size_t off = 0; void* p = nullptr;
  if (napi_get_typedarray_info(env, v, &type, len, &p, &ab, &off) != napi_ok || type != napi_uint8_array) {
    napi_throw_type_error(env, nullptr, "hzwitness: Uint8Array expected");
    return false;
  }
  *data = static_cast<uint8_t*>(p);
  return true;
}
static void hz_ctx_finalize(napi_env, void* data, void*) { hzw_destroy(static_cast<hzw_ctx*>(data)); }

static napi_value Create(napi_env env, napi_callback_info info) {
  size_t argc = 8; napi_value a[8];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  if (argc < 5) { napi_throw_type_error(env, nullptr, "create(kind, nTx, nLevels, maxL1Tx, maxFeeTx[, device, nGpus, symPath])"); return nullptr; }
  int32_t v[7] = {0, 0, 0, 0, 0, 0, 1};
  for (size_t i = 0; i < 7 && i < argc; i++) NAPI_OK(napi_get_value_int32(env, a[i], &v[i]));
  hzw_params p;
  memset(&p, 0, sizeof p);
  p.circuit_kind = v[0]; p.nTx = v[1]; p.nLevels = v[2]; p.maxL1Tx = v[3]; p.maxFeeTx = v[4]; p.device = v[5]; p.n_gpus = v[6]; p.flags = 0;
  std::string sym;
  bool have_sym = false;
  if (argc >= 8) {
    napi_valuetype t;
    NAPI_OK(napi_typeof(env, a[7], &t));
    if (t == napi_string) { if (!hz_string_arg(env, a[7], sym)) return nullptr; have_sym = true; }
  }
  hzw_ctx* ctx = nullptr;
  int st = hzw_create(&p, have_sym ? sym.c_str() : nullptr, &ctx);
  if (st != HZW_OK) {
    napi_throw_error(env, nullptr, st == HZW_E_DEVICE ? "hzw_create: no usable CUDA device (there is no CPU fallback)" : "hzw_create: bad parameters or layout map");
    return nullptr;
  }
  napi_value ext;
  NAPI_OK(napi_create_external(env, ctx, hz_ctx_finalize, nullptr, &ext));
  return ext;
}

static napi_value NVars(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value a[1];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  hzw_ctx* ctx = hz_ctx_arg(env, a[0]);
  if (!ctx) return nullptr;
  napi_value r;
  NAPI_OK(napi_create_double(env, (double)hzw_nvars(ctx), &r));
  return r;
}
static napi_value NInputs(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value a[1];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  hzw_ctx* ctx = hz_ctx_arg(env, a[0]);
  if (!ctx) return nullptr;
  napi_value r;
  NAPI_OK(napi_create_double(env, (double)hzw_ninputs(ctx), &r));
  return r;
}
static napi_value InputIndex(napi_env env, napi_callback_info info) {
  size_t argc = 2; napi_value a[2];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  hzw_ctx* ctx = hz_ctx_arg(env, a[0]);
  std::string name;
  if (!ctx || !hz_string_arg(env, a[1], name)) return nullptr;
  uint64_t off = 0, len = 0;
  if (hzw_input_index(ctx, name.c_str(), &off, &len) != HZW_OK) {
    napi_throw_error(env, nullptr, ("Signal " + name + " is not an input of the circuit.").c_str());
    return nullptr;
  }
  napi_value arr, vo, vl;
  NAPI_OK(napi_create_array_with_length(env, 2, &arr));
  NAPI_OK(napi_create_double(env, (double)off, &vo));
  NAPI_OK(napi_create_double(env, (double)len, &vl));
  NAPI_OK(napi_set_element(env, arr, 0, vo));
  NAPI_OK(napi_set_element(env, arr, 1, vl));
  return arr;
}

// ---- calculate: hzw_calculate on a worker thread, the witness resolved as a Uint8Array --------------------------
struct HzJob {
  hzw_ctx* ctx;
  uint8_t* in; uint8_t* out;
  int sanity, st;
  hzw_error err;
  napi_deferred deferred;
  napi_async_work work;
  napi_ref in_ref, out_ref;      // keep the two typed arrays alive while the worker runs
};
static void HzExecute(napi_env, void* data) {
  HzJob* j = static_cast<HzJob*>(data);
  j->st = hzw_calculate(j->ctx, j->in, j->sanity, j->out, &j->err);
}
static void HzComplete(napi_env env, napi_status status, void* data) {
  HzJob* j = static_cast<HzJob*>(data);
  if (status == napi_ok && j->st == HZW_OK) {
    napi_value out;
    napi_get_reference_value(env, j->out_ref, &out);
    napi_resolve_deferred(env, j->deferred, out);
  } else {
    const std::string m = status != napi_ok ? std::string("hzwitness: worker cancelled") : hz_error_message(j->st, j->err);
    napi_value msg, err;
    napi_create_string_utf8(env, m.c_str(), NAPI_AUTO_LENGTH, &msg);
    napi_create_error(env, nullptr, msg, &err);
    napi_reject_deferred(env, j->deferred, err);
  }
  napi_delete_reference(env, j->in_ref);
  napi_delete_reference(env, j->out_ref);
  napi_delete_async_work(env, j->work);
  delete j;
}
static napi_value Calculate(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value a[3];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  hzw_ctx* ctx = hz_ctx_arg(env, a[0]);
  if (!ctx) return nullptr;
  uint8_t* in = nullptr; size_t in_len = 0;
  if (!hz_bytes_arg(env, a[1], &in, &in_len)) return nullptr;
  if (in_len != 32 * (size_t)hzw_ninputs(ctx)) { napi_throw_error(env, nullptr, "Not all inputs have been set."); return nullptr; }
  bool sanity = false;
  if (argc >= 3) NAPI_OK(napi_get_value_bool(env, a[2], &sanity));
  const size_t out_len = 32 * (size_t)hzw_nvars(ctx);
  void* out_data = nullptr;
  napi_value ab, out;
  NAPI_OK(napi_create_arraybuffer(env, out_len, &out_data, &ab));
  NAPI_OK(napi_create_typedarray(env, napi_uint8_array, out_len, ab, 0, &out));
  HzJob* j = new HzJob();
  memset(j, 0, sizeof *j);
  j->ctx = ctx; j->in = in; j->out = static_cast<uint8_t*>(out_data); j->sanity = sanity ? 1 : 0;
  napi_value promise, name;
  if (napi_create_promise(env, &j->deferred, &promise) != napi_ok || napi_create_reference(env, a[1], 1, &j->in_ref) != napi_ok ||
      napi_create_reference(env, out, 1, &j->out_ref) != napi_ok ||
      napi_create_string_utf8(env, "hzw_calculate", NAPI_AUTO_LENGTH, &name) != napi_ok ||
      napi_create_async_work(env, nullptr, name, HzExecute, HzComplete, j, &j->work) != napi_ok ||
      napi_queue_async_work(env, j->work) != napi_ok) {
    delete j;
    napi_throw_error(env, nullptr, "hzwitness: could not queue the calculation");
    return nullptr;
  }
  return promise;
}

static napi_value GetSignal(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value a[3];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  hzw_ctx* ctx = hz_ctx_arg(env, a[0]);
  uint8_t* w = nullptr; size_t w_len = 0; std::string name;
  if (!ctx || !hz_bytes_arg(env, a[1], &w, &w_len) || !hz_string_arg(env, a[2], name)) return nullptr;
  if (w_len != 32 * (size_t)hzw_nvars(ctx)) { napi_throw_error(env, nullptr, "hzwitness: witness of the wrong size"); return nullptr; }
  void* out_data = nullptr;
  napi_value ab, out;
  NAPI_OK(napi_create_arraybuffer(env, 32, &out_data, &ab));
  if (hzw_get_signal(ctx, w, name.c_str(), static_cast<uint8_t*>(out_data)) != HZW_OK) {
    napi_throw_error(env, nullptr, ("Signal not found: " + name).c_str());
    return nullptr;
  }
  NAPI_OK(napi_create_typedarray(env, napi_uint8_array, 32, ab, 0, &out));
  return out;
}
static napi_value SignalIndex(napi_env env, napi_callback_info info) {
  size_t argc = 2; napi_value a[2];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  hzw_ctx* ctx = hz_ctx_arg(env, a[0]);
  std::string name;
  if (!ctx || !hz_string_arg(env, a[1], name)) return nullptr;
  uint64_t idx = 0;
  if (hzw_signal_index(ctx, name.c_str(), &idx) != HZW_OK) { napi_throw_error(env, nullptr, ("Signal not found: " + name).c_str()); return nullptr; }
  napi_value r;
  NAPI_OK(napi_create_double(env, (double)idx, &r));
  return r;
}
static napi_value WriteWtns(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value a[3];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  hzw_ctx* ctx = hz_ctx_arg(env, a[0]);
  uint8_t* w = nullptr; size_t w_len = 0; std::string path;
  if (!ctx || !hz_bytes_arg(env, a[1], &w, &w_len) || !hz_string_arg(env, a[2], path)) return nullptr;
  if (w_len != 32 * (size_t)hzw_nvars(ctx)) { napi_throw_error(env, nullptr, "hzwitness: witness of the wrong size"); return nullptr; }
  if (hzw_write_wtns(ctx, w, path.c_str()) != HZW_OK) { napi_throw_error(env, nullptr, ("cannot write " + path).c_str()); return nullptr; }
  napi_value undef;
  NAPI_OK(napi_get_undefined(env, &undef));
  return undef;
}
static napi_value CheckConstraints(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value a[3];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  hzw_ctx* ctx = hz_ctx_arg(env, a[0]);
  std::string path;
  if (!ctx || !hz_string_arg(env, a[1], path)) return nullptr;
  uint8_t* w = nullptr; size_t w_len = 0;
  if (argc >= 3) {
    napi_valuetype t;
    NAPI_OK(napi_typeof(env, a[2], &t));
    if (t != napi_null && t != napi_undefined && !hz_bytes_arg(env, a[2], &w, &w_len)) return nullptr;
  }
  uint64_t n = 0, bad = 0, first = 0;
  if (hzw_r1cs_check(ctx, path.c_str(), w, &n, &bad, &first, nullptr) != HZW_OK) { napi_throw_error(env, nullptr, ("cannot check " + path).c_str()); return nullptr; }
  napi_value arr, v0, v1, v2;
  NAPI_OK(napi_create_array_with_length(env, 3, &arr));
  NAPI_OK(napi_create_double(env, (double)n, &v0));
  NAPI_OK(napi_create_double(env, (double)bad, &v1));
  NAPI_OK(napi_create_double(env, bad ? (double)first : -1.0, &v2));
  NAPI_OK(napi_set_element(env, arr, 0, v0));
  NAPI_OK(napi_set_element(env, arr, 1, v1));
  NAPI_OK(napi_set_element(env, arr, 2, v2));
  return arr;
}

static napi_value Init(napi_env env, napi_value exports) {
  const napi_property_descriptor props[] = {
      {"create", nullptr, Create, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"nvars", nullptr, NVars, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"ninputs", nullptr, NInputs, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"inputIndex", nullptr, InputIndex, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"calculate", nullptr, Calculate, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"getSignal", nullptr, GetSignal, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"signalIndex", nullptr, SignalIndex, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"writeWtns", nullptr, WriteWtns, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"checkConstraints", nullptr, CheckConstraints, nullptr, nullptr, nullptr, napi_default, nullptr},
  };
  NAPI_OK(napi_define_properties(env, exports, sizeof props / sizeof props[0], props));
  return exports;
}
NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
