// Test stub, not Node's header: just the Node-API declarations bindings/node/zsgpu_addon.cc uses, so that
// tests/test_bindings_compile.py can type-check the addon against include/zsgpu.h in an image without node.
#include <cstddef>
#include <cstdint>
typedef struct napi_env__* napi_env; typedef struct napi_value__* napi_value; typedef struct napi_callback_info__* napi_callback_info;
typedef enum { napi_ok } napi_status;
typedef enum { napi_int8_array, napi_uint8_array, napi_biguint64_array = 10 } napi_typedarray_type;
typedef napi_value (*napi_callback)(napi_env, napi_callback_info);
typedef void (*napi_finalize)(napi_env, void*, void*);
#define NAPI_AUTO_LENGTH SIZE_MAX
#define NAPI_MODULE(n, f) extern "C" napi_value napi_register_##n(napi_env e, napi_value x) { return f(e, x); }
extern "C" {
napi_status napi_create_array_with_length(napi_env, size_t, napi_value*);
napi_status napi_create_buffer_copy(napi_env, size_t, const void*, void**, napi_value*);
napi_status napi_create_double(napi_env, double, napi_value*);
napi_status napi_create_external(napi_env, void*, napi_finalize, void*, napi_value*);
napi_status napi_create_function(napi_env, const char*, size_t, napi_callback, void*, napi_value*);
napi_status napi_create_object(napi_env, napi_value*);
napi_status napi_get_cb_info(napi_env, napi_callback_info, size_t*, napi_value*, napi_value*, void**);
napi_status napi_get_typedarray_info(napi_env, napi_value, napi_typedarray_type*, size_t*, void**, napi_value*, size_t*);
napi_status napi_get_value_external(napi_env, napi_value, void**);
napi_status napi_get_value_int32(napi_env, napi_value, int32_t*);
napi_status napi_get_value_uint32(napi_env, napi_value, uint32_t*);
napi_status napi_set_element(napi_env, napi_value, uint32_t, napi_value);
napi_status napi_set_named_property(napi_env, napi_value, const char*, napi_value);
napi_status napi_throw_error(napi_env, const char*, const char*);
}
