/* Declarations of the N-API subset hzwitness_napi.cc uses, for a syntax / type check where Node's headers are absent
 * (`g++ -fsyntax-only -I test -include test/node_api_stub.h`, run by tests/test_bindings.py).  NOT a Node header. */
#ifndef NODE_API_STUB_H
#define NODE_API_STUB_H
#include <stddef.h>
#include <stdint.h>
typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_ref__* napi_ref;
typedef struct napi_deferred__* napi_deferred;
typedef struct napi_callback_info__* napi_callback_info;
typedef struct napi_async_work__* napi_async_work;
typedef enum { napi_ok, napi_invalid_arg, napi_generic_failure, napi_cancelled } napi_status;
typedef enum { napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object, napi_function, napi_external, napi_bigint } napi_valuetype;
typedef enum { napi_int8_array, napi_uint8_array, napi_uint8_clamped_array } napi_typedarray_type;
typedef enum { napi_default = 0 } napi_property_attributes;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void* finalize_data, void* finalize_hint);
typedef void (*napi_async_execute_callback)(napi_env env, void* data);
typedef void (*napi_async_complete_callback)(napi_env env, napi_status status, void* data);
typedef struct { const char* utf8name; napi_value name; napi_callback method; napi_callback getter; napi_callback setter; napi_value value; napi_property_attributes attributes; void* data; } napi_property_descriptor;
#define NAPI_AUTO_LENGTH ((size_t)-1)
#define NAPI_MODULE(name, init) extern "C" napi_value napi_register_module_v1(napi_env env, napi_value exports) { return init(env, exports); }
#define NODE_GYP_MODULE_NAME hzwitness_napi
extern "C" {
napi_status napi_throw_error(napi_env, const char*, const char*);
napi_status napi_throw_type_error(napi_env, const char*, const char*);
napi_status napi_get_cb_info(napi_env, napi_callback_info, size_t*, napi_value*, napi_value*, void**);
napi_status napi_get_value_int32(napi_env, napi_value, int32_t*);
napi_status napi_get_value_bool(napi_env, napi_value, bool*);
napi_status napi_get_value_external(napi_env, napi_value, void**);
napi_status napi_get_value_string_utf8(napi_env, napi_value, char*, size_t, size_t*);
napi_status napi_typeof(napi_env, napi_value, napi_valuetype*);
napi_status napi_is_typedarray(napi_env, napi_value, bool*);
napi_status napi_get_typedarray_info(napi_env, napi_value, napi_typedarray_type*, size_t*, void**, napi_value*, size_t*);
napi_status napi_create_external(napi_env, void*, napi_finalize, void*, napi_value*);
napi_status napi_create_double(napi_env, double, napi_value*);
napi_status napi_create_array_with_length(napi_env, size_t, napi_value*);
napi_status napi_set_element(napi_env, napi_value, uint32_t, napi_value);
napi_status napi_create_arraybuffer(napi_env, size_t, void**, napi_value*);
napi_status napi_create_typedarray(napi_env, napi_typedarray_type, size_t, napi_value, size_t, napi_value*);
napi_status napi_create_promise(napi_env, napi_deferred*, napi_value*);
napi_status napi_resolve_deferred(napi_env, napi_deferred, napi_value);
napi_status napi_reject_deferred(napi_env, napi_deferred, napi_value);
napi_status napi_create_reference(napi_env, napi_value, uint32_t, napi_ref*);
napi_status napi_delete_reference(napi_env, napi_ref);
napi_status napi_get_reference_value(napi_env, napi_ref, napi_value*);
napi_status napi_create_string_utf8(napi_env, const char*, size_t, napi_value*);
napi_status napi_create_error(napi_env, napi_value, napi_value, napi_value*);
napi_status napi_create_async_work(napi_env, napi_value, napi_value, napi_async_execute_callback, napi_async_complete_callback, void*, napi_async_work*);
napi_status napi_queue_async_work(napi_env, napi_async_work);
napi_status napi_delete_async_work(napi_env, napi_async_work);
napi_status napi_get_undefined(napi_env, napi_value*);
napi_status napi_define_properties(napi_env, napi_value, size_t, const napi_property_descriptor*);
}
#endif
