/* Minimal stand-in for Node's <node_api.h>, ONLY so that integration/rmb_napi.c can be syntax- and
 * type-checked in an image without Node.js (tests/test_integration_cpu.py runs `gcc -fsyntax-only`).
 * Declarations follow the stable N-API (version 8) signatures for the calls the addon uses. */
#ifndef RMB_STUB_NODE_API_H_
#define RMB_STUB_NODE_API_H_
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_callback_info__* napi_callback_info;
typedef enum { napi_ok, napi_invalid_arg, napi_generic_failure } napi_status;
typedef enum { napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object, napi_function, napi_external, napi_bigint } napi_valuetype;
typedef enum { napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array, napi_int32_array,
               napi_uint32_array, napi_float32_array, napi_float64_array, napi_bigint64_array, napi_biguint64_array } napi_typedarray_type;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void* finalize_data, void* finalize_hint);
#define NAPI_AUTO_LENGTH SIZE_MAX
napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_get_value_external(napi_env env, napi_value value, void** result);
napi_status napi_get_value_int32(napi_env env, napi_value value, int32_t* result);
napi_status napi_get_value_uint32(napi_env env, napi_value value, uint32_t* result);
napi_status napi_get_value_double(napi_env env, napi_value value, double* result);
napi_status napi_get_value_string_utf8(napi_env env, napi_value value, char* buf, size_t bufsize, size_t* result);
napi_status napi_create_int32(napi_env env, int32_t value, napi_value* result);
napi_status napi_create_external(napi_env env, void* data, napi_finalize finalize_cb, void* finalize_hint, napi_value* result);
napi_status napi_create_object(napi_env env, napi_value* result);
napi_status napi_create_string_utf8(napi_env env, const char* str, size_t length, napi_value* result);
napi_status napi_create_function(napi_env env, const char* utf8name, size_t length, napi_callback cb, void* data, napi_value* result);
napi_status napi_get_undefined(napi_env env, napi_value* result);
napi_status napi_typeof(napi_env env, napi_value value, napi_valuetype* result);
napi_status napi_get_array_length(napi_env env, napi_value value, uint32_t* result);
napi_status napi_get_element(napi_env env, napi_value object, uint32_t index, napi_value* result);
napi_status napi_get_named_property(napi_env env, napi_value object, const char* utf8name, napi_value* result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char* utf8name, napi_value value);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type* type, size_t* length, void** data,
                                     napi_value* arraybuffer, size_t* byte_offset);
#define NODE_GYP_MODULE_NAME rmb
#define NAPI_MODULE(modname, regfunc) napi_value napi_register_module_v1(napi_env env, napi_value exports) { return regfunc(env, exports); }
#endif
