/* rmb_napi.c -- N-API addon mapping include/rmb.h 1:1 into JavaScript.
 *
 * NOT BUILT IN THIS IMAGE: there is no Node.js, node_api.h or node-gyp here.  The
 * same C symbols are exercised through Python ctypes by tests/; this file is the binding a
 * maintainer of radian628/raymarching-engine would compile with
 *     cc -shared -fPIC -I$(node -p "process.execPath+'/../../include/node'") -I../include \
 *        rmb_napi.c -L../raymarching_engine_b200 -lraymarch_b200 -o rmb.node
 * Handles travel as napi externals; buffers as ArrayBuffer / TypedArray (no copies beyond the
 * readback itself).  Errors are values (status codes + strings), never JS exceptions, like the
 * reference's ShaderError (client/src/renderer/ShaderCache.tsx:8-11).
 */
#include <node_api.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "rmb.h"

#define ARGS(n)                                                    \
    size_t argc = (n);                                             \
    napi_value argv[(n) > 0 ? (n) : 1];                            \
    napi_get_cb_info(env, info, &argc, argv, NULL, NULL)

static void* ext(napi_env env, napi_value v) { void* p = NULL; napi_get_value_external(env, v, &p); return p; }
static int32_t i32(napi_env env, napi_value v) { int32_t x = 0; napi_get_value_int32(env, v, &x); return x; }
static double f64(napi_env env, napi_value v) { double x = 0; napi_get_value_double(env, v, &x); return x; }
static napi_value mk_i32(napi_env env, int32_t x) { napi_value v; napi_create_int32(env, x, &v); return v; }
static napi_value mk_ext(napi_env env, void* p) {
    napi_value v;
    if (!p) { napi_get_undefined(env, &v); return v; }   /* NULL handle -> undefined, like the reference */
    napi_create_external(env, p, NULL, NULL, &v);
    return v;
}
static char* str(napi_env env, napi_value v, size_t* len) {
    size_t n = 0;
    napi_get_value_string_utf8(env, v, NULL, 0, &n);
    char* s = (char*)malloc(n + 1);
    napi_get_value_string_utf8(env, v, s, n + 1, &n);
    if (len) *len = n;
    return s;
}

/* ctxCreate(device, rank, nRanks, tileRows) -> external | undefined      LoadRenderJobContext.tsx:268-287 */
static napi_value CtxCreate(napi_env env, napi_callback_info info) {
    ARGS(4);
    return mk_ext(env, rmb_ctx_create(i32(env, argv[0]), i32(env, argv[1]), i32(env, argv[2]), i32(env, argv[3])));
}
static napi_value CtxDestroy(napi_env env, napi_callback_info info) { ARGS(1); rmb_ctx_destroy((rmb_ctx*)ext(env, argv[0])); return NULL; }
static napi_value LastError(napi_env env, napi_callback_info info) {
    ARGS(1);
    napi_valuetype t; napi_typeof(env, argv[0], &t);
    const char* s = rmb_last_error(t == napi_external ? (rmb_ctx*)ext(env, argv[0]) : NULL);
    napi_value v; napi_create_string_utf8(env, s ? s : "", NAPI_AUTO_LENGTH, &v); return v;
}

/* programGet(ctx, sceneGlsl, flavour, spec[{name,type,data:number[]}]) ->
 *   {program: external} | {type: "fragment"|"program"|"general", infoLog}      ShaderCache.tsx:91-119 */
typedef struct { rmb_spec_uniform* spec; char** names; uint32_t n; } spec_list;
static spec_list parse_spec(napi_env env, napi_value arr) {
    spec_list l; l.n = 0;
    napi_get_array_length(env, arr, &l.n);
    l.spec = (rmb_spec_uniform*)calloc(l.n ? l.n : 1, sizeof *l.spec);
    l.names = (char**)calloc(l.n ? l.n : 1, sizeof *l.names);
    for (uint32_t k = 0; k < l.n; k++) {
        napi_value e, f; napi_get_element(env, arr, k, &e);
        napi_get_named_property(env, e, "name", &f); l.names[k] = str(env, f, NULL); l.spec[k].name = l.names[k];
        napi_get_named_property(env, e, "type", &f); l.spec[k].type = i32(env, f);
        napi_get_named_property(env, e, "data", &f);
        uint32_t c = 0; napi_get_array_length(env, f, &c); if (c > 4) c = 4; l.spec[k].count = (int)c;
        for (uint32_t j = 0; j < c; j++) {
            napi_value x; napi_get_element(env, f, j, &x);
            if (l.spec[k].type == RMB_UNIFORM_F) l.spec[k].data.f[j] = (float)f64(env, x);
            else if (l.spec[k].type == RMB_UNIFORM_I) l.spec[k].data.i[j] = i32(env, x);
            else { uint32_t u = 0; napi_get_value_uint32(env, x, &u); l.spec[k].data.u[j] = u; }
        }
    }
    return l;
}
static void free_spec(spec_list l) {
    for (uint32_t k = 0; k < l.n; k++) free(l.names[k]);
    free(l.names); free(l.spec);
}
/* {program: external} on success, else the reference's ShaderError value {type, infoLog} */
static napi_value program_result(napi_env env, rmb_status st, void* prog, const char* etype, const char* log) {
    napi_value out; napi_create_object(env, &out);
    if (st == RMB_OK) napi_set_named_property(env, out, "program", mk_ext(env, prog));
    else {
        napi_value a, b; napi_create_string_utf8(env, etype[0] ? etype : "general", NAPI_AUTO_LENGTH, &a);
        napi_create_string_utf8(env, log, NAPI_AUTO_LENGTH, &b);
        napi_set_named_property(env, out, "type", a); napi_set_named_property(env, out, "infoLog", b);
    }
    return out;
}
static napi_value ProgramGet(napi_env env, napi_callback_info info) {
    ARGS(4);
    size_t len = 0; char* src = str(env, argv[1], &len);
    spec_list l = parse_spec(env, argv[3]);
    rmb_program* prog = NULL; char etype[16] = {0}; char* log = (char*)calloc(1, 1 << 16);
    rmb_status st = rmb_program_get((rmb_ctx*)ext(env, argv[0]), src, len, i32(env, argv[2]), l.spec, (int)l.n, &prog, etype, log, 1 << 16);
    napi_value out = program_result(env, st, prog, etype, log);
    free_spec(l); free(src); free(log);
    return out;
}

/* uniformSet(program, name, type, count, TypedArray)                         Uniforms.tsx:34-46
 * uniformSetArray(program, name, type, components, nElements, TypedArray)    RenderJobExecutor.tsx:268-291
 * uniformMatrix4(program, name, Float32Array(16))                            RenderJobExecutor.tsx:293-297 */
static void* typed(napi_env env, napi_value v) {
    void* data = NULL; size_t n; napi_typedarray_type t; napi_value ab; size_t off;
    napi_get_typedarray_info(env, v, &t, &n, &data, &ab, &off);
    return data;
}
static napi_value UniformSet(napi_env env, napi_callback_info info) {
    ARGS(5); char* name = str(env, argv[1], NULL);
    rmb_status st = rmb_uniform_set((rmb_program*)ext(env, argv[0]), name, i32(env, argv[2]), i32(env, argv[3]), typed(env, argv[4]));
    free(name); return mk_i32(env, st);
}
static napi_value UniformSetArray(napi_env env, napi_callback_info info) {
    ARGS(6); char* name = str(env, argv[1], NULL);
    rmb_status st = rmb_uniform_set_array((rmb_program*)ext(env, argv[0]), name, i32(env, argv[2]), i32(env, argv[3]), i32(env, argv[4]), typed(env, argv[5]));
    free(name); return mk_i32(env, st);
}
static napi_value UniformMatrix4(napi_env env, napi_callback_info info) {
    ARGS(3); char* name = str(env, argv[1], NULL);
    rmb_status st = rmb_uniform_matrix4((rmb_program*)ext(env, argv[0]), name, (const float*)typed(env, argv[2]));
    free(name); return mk_i32(env, st);
}

/* uniformsSetFrame(program, TypedArray over an rmb_frame_uniforms block): the whole per-sample upload of
 * RenderJobExecutor.tsx:212-297 in one crossing (include/rmb.h; a DataView / Float32Array + Int32Array over one
 * ArrayBuffer of sizeof(rmb_frame_uniforms) bytes, fields in declaration order) */
static napi_value UniformsSetFrame(napi_env env, napi_callback_info info) {
    ARGS(2);
    return mk_i32(env, rmb_uniforms_set_frame((rmb_program*)ext(env, argv[0]), (const rmb_frame_uniforms*)typed(env, argv[1])));
}

/* fbAcquire(ctx, w, h, frameid) -> external | undefined; fbRelease(ctx, w, h, frameid)   LoadRenderJobContext.tsx:184-249 */
static napi_value FbAcquire(napi_env env, napi_callback_info info) {
    ARGS(4);
    return mk_ext(env, rmb_fb_acquire((rmb_ctx*)ext(env, argv[0]), i32(env, argv[1]), i32(env, argv[2]), (int64_t)f64(env, argv[3])));
}
static napi_value FbRelease(napi_env env, napi_callback_info info) {
    ARGS(4); rmb_fb_release((rmb_ctx*)ext(env, argv[0]), i32(env, argv[1]), i32(env, argv[2]), (int64_t)f64(env, argv[3])); return NULL;
}
static napi_value FbLocalRows(napi_env env, napi_callback_info info) { ARGS(1); return mk_i32(env, rmb_fb_local_rows((rmb_fb*)ext(env, argv[0]))); }

/* renderSample(ctx, program, fb, x, y, w, h)     gl.scissor + drawArrays + blit, RenderJobExecutor.tsx:181-326 */
static napi_value RenderSample(napi_env env, napi_callback_info info) {
    ARGS(7);
    return mk_i32(env, rmb_render_sample((rmb_ctx*)ext(env, argv[0]), (rmb_program*)ext(env, argv[1]), (rmb_fb*)ext(env, argv[2]),
                                         i32(env, argv[3]), i32(env, argv[4]), i32(env, argv[5]), i32(env, argv[6])));
}
/* present(ctx, fb, brightness, Uint8Array rgba8, Float32Array depth | null)   index.tsx:25-59, :470-476 */
static napi_value Present(napi_env env, napi_callback_info info) {
    ARGS(5);
    napi_valuetype t; napi_typeof(env, argv[4], &t);
    return mk_i32(env, rmb_present((rmb_ctx*)ext(env, argv[0]), (rmb_fb*)ext(env, argv[1]), (float)f64(env, argv[2]),
                                   (uint8_t*)typed(env, argv[3]), t == napi_object ? (float*)typed(env, argv[4]) : NULL));
}

/* presentAsync(ctx, fb, brightness, Uint8Array rgba8, Float32Array depth | null) / presentWait(ctx, fb): the
 * non-blocking pair (WebGL draws return immediately; only the capture waits, index.tsx:470-476) */
static napi_value PresentAsync(napi_env env, napi_callback_info info) {
    ARGS(5);
    napi_valuetype t; napi_typeof(env, argv[4], &t);
    return mk_i32(env, rmb_present_async((rmb_ctx*)ext(env, argv[0]), (rmb_fb*)ext(env, argv[1]), (float)f64(env, argv[2]),
                                         (uint8_t*)typed(env, argv[3]), t == napi_object ? (float*)typed(env, argv[4]) : NULL));
}
static napi_value PresentWait(napi_env env, napi_callback_info info) {
    ARGS(2);
    return mk_i32(env, rmb_present_wait((rmb_ctx*)ext(env, argv[0]), (rmb_fb*)ext(env, argv[1])));
}
static napi_value Sync(napi_env env, napi_callback_info info) { ARGS(1); return mk_i32(env, rmb_sync((rmb_ctx*)ext(env, argv[0]))); }

/* ---- device groups (include/rmb.h rmb_group_*): every GPU of the box behind one handle, so the single-threaded
 * host of index.tsx:236-263 renders one frame on all of them; same call shapes as above with `group` for `ctx`.
 * groupCreate(devices: number[], tileRows) -> external | undefined */
static napi_value GroupCreate(napi_env env, napi_callback_info info) {
    ARGS(2);
    uint32_t n = 0; napi_get_array_length(env, argv[0], &n);
    int* dev = (int*)calloc(n ? n : 1, sizeof *dev);
    for (uint32_t k = 0; k < n; k++) { napi_value e; napi_get_element(env, argv[0], k, &e); dev[k] = i32(env, e); }
    rmb_group* g = rmb_group_create(dev, (int)n, i32(env, argv[1]));
    free(dev);
    return mk_ext(env, g);
}
static napi_value GroupDestroy(napi_env env, napi_callback_info info) { ARGS(1); rmb_group_destroy((rmb_group*)ext(env, argv[0])); return NULL; }
static napi_value GroupLastError(napi_env env, napi_callback_info info) {
    ARGS(1);
    napi_valuetype t; napi_typeof(env, argv[0], &t);
    const char* s = rmb_group_last_error(t == napi_external ? (rmb_group*)ext(env, argv[0]) : NULL);
    napi_value v; napi_create_string_utf8(env, s ? s : "", NAPI_AUTO_LENGTH, &v); return v;
}
static napi_value GroupSize(napi_env env, napi_callback_info info) { ARGS(1); return mk_i32(env, rmb_group_size((rmb_group*)ext(env, argv[0]))); }
static napi_value GroupSync(napi_env env, napi_callback_info info) { ARGS(1); return mk_i32(env, rmb_group_sync((rmb_group*)ext(env, argv[0]))); }
static napi_value GroupProgramGet(napi_env env, napi_callback_info info) {
    ARGS(4);
    size_t len = 0; char* src = str(env, argv[1], &len);
    spec_list l = parse_spec(env, argv[3]);
    rmb_group_program* prog = NULL; char etype[16] = {0}; char* log = (char*)calloc(1, 1 << 16);
    rmb_status st = rmb_group_program_get((rmb_group*)ext(env, argv[0]), src, len, i32(env, argv[2]), l.spec, (int)l.n, &prog, etype, log, 1 << 16);
    napi_value out = program_result(env, st, prog, etype, log);
    free_spec(l); free(src); free(log);
    return out;
}
static napi_value GroupUniformSet(napi_env env, napi_callback_info info) {
    ARGS(5); char* name = str(env, argv[1], NULL);
    rmb_status st = rmb_group_uniform_set((rmb_group_program*)ext(env, argv[0]), name, i32(env, argv[2]), i32(env, argv[3]), typed(env, argv[4]));
    free(name); return mk_i32(env, st);
}
static napi_value GroupUniformSetArray(napi_env env, napi_callback_info info) {
    ARGS(6); char* name = str(env, argv[1], NULL);
    rmb_status st = rmb_group_uniform_set_array((rmb_group_program*)ext(env, argv[0]), name, i32(env, argv[2]), i32(env, argv[3]), i32(env, argv[4]), typed(env, argv[5]));
    free(name); return mk_i32(env, st);
}
static napi_value GroupUniformMatrix4(napi_env env, napi_callback_info info) {
    ARGS(3); char* name = str(env, argv[1], NULL);
    rmb_status st = rmb_group_uniform_matrix4((rmb_group_program*)ext(env, argv[0]), name, (const float*)typed(env, argv[2]));
    free(name); return mk_i32(env, st);
}
static napi_value GroupUniformsSetFrame(napi_env env, napi_callback_info info) {
    ARGS(2);
    return mk_i32(env, rmb_group_uniforms_set_frame((rmb_group_program*)ext(env, argv[0]), (const rmb_frame_uniforms*)typed(env, argv[1])));
}
static napi_value GroupFbAcquire(napi_env env, napi_callback_info info) {
    ARGS(4);
    return mk_ext(env, rmb_group_fb_acquire((rmb_group*)ext(env, argv[0]), i32(env, argv[1]), i32(env, argv[2]), (int64_t)f64(env, argv[3])));
}
static napi_value GroupFbRelease(napi_env env, napi_callback_info info) {
    ARGS(4); rmb_group_fb_release((rmb_group*)ext(env, argv[0]), i32(env, argv[1]), i32(env, argv[2]), (int64_t)f64(env, argv[3])); return NULL;
}
static napi_value GroupRenderSample(napi_env env, napi_callback_info info) {
    ARGS(7);
    return mk_i32(env, rmb_group_render_sample((rmb_group*)ext(env, argv[0]), (rmb_group_program*)ext(env, argv[1]), (rmb_group_fb*)ext(env, argv[2]),
                                               i32(env, argv[3]), i32(env, argv[4]), i32(env, argv[5]), i32(env, argv[6])));
}
/* groupPresent(group, fb, brightness, Uint8Array rgba8 (whole frame), Float32Array depth | null) */
static napi_value GroupPresent(napi_env env, napi_callback_info info) {
    ARGS(5);
    napi_valuetype t; napi_typeof(env, argv[4], &t);
    return mk_i32(env, rmb_group_present((rmb_group*)ext(env, argv[0]), (rmb_group_fb*)ext(env, argv[1]), (float)f64(env, argv[2]),
                                         (uint8_t*)typed(env, argv[3]), t == napi_object ? (float*)typed(env, argv[4]) : NULL));
}
/* multi-process alternative (one context per process, frame assembled in rank 0's buffer over CUDA IPC):
 * setGatherTarget(ctx, external | null, bytes), fbScatterRows(ctx, fb, which, external), displayPlanes(ctx, color, nd, out, w, h, brightness) */
static napi_value SetGatherTarget(napi_env env, napi_callback_info info) {
    ARGS(3);
    napi_valuetype t; napi_typeof(env, argv[1], &t);
    return mk_i32(env, rmb_ctx_set_gather_target((rmb_ctx*)ext(env, argv[0]), t == napi_external ? ext(env, argv[1]) : NULL, (size_t)f64(env, argv[2])));
}
static napi_value FbScatterRows(napi_env env, napi_callback_info info) {
    ARGS(4);
    return mk_i32(env, rmb_fb_scatter_rows((rmb_ctx*)ext(env, argv[0]), (rmb_fb*)ext(env, argv[1]), i32(env, argv[2]), ext(env, argv[3])));
}
static napi_value DisplayPlanes(napi_env env, napi_callback_info info) {
    ARGS(7);
    return mk_i32(env, rmb_display_planes((rmb_ctx*)ext(env, argv[0]), ext(env, argv[1]), ext(env, argv[2]), ext(env, argv[3]),
                                          i32(env, argv[4]), i32(env, argv[5]), (float)f64(env, argv[6])));
}

#define EXPORT(name, fn) do { napi_value f; napi_create_function(env, name, NAPI_AUTO_LENGTH, fn, NULL, &f); napi_set_named_property(env, exports, name, f); } while (0)
static napi_value Init(napi_env env, napi_value exports) {
    EXPORT("ctxCreate", CtxCreate); EXPORT("ctxDestroy", CtxDestroy); EXPORT("lastError", LastError);
    EXPORT("programGet", ProgramGet); EXPORT("uniformSet", UniformSet); EXPORT("uniformSetArray", UniformSetArray);
    EXPORT("uniformMatrix4", UniformMatrix4); EXPORT("uniformsSetFrame", UniformsSetFrame); EXPORT("fbAcquire", FbAcquire); EXPORT("fbRelease", FbRelease);
    EXPORT("fbLocalRows", FbLocalRows); EXPORT("renderSample", RenderSample); EXPORT("present", Present);
    EXPORT("presentAsync", PresentAsync); EXPORT("presentWait", PresentWait); EXPORT("sync", Sync);
    EXPORT("groupCreate", GroupCreate); EXPORT("groupDestroy", GroupDestroy); EXPORT("groupLastError", GroupLastError);
    EXPORT("groupSize", GroupSize); EXPORT("groupSync", GroupSync); EXPORT("groupProgramGet", GroupProgramGet);
    EXPORT("groupUniformSet", GroupUniformSet); EXPORT("groupUniformSetArray", GroupUniformSetArray);
    EXPORT("groupUniformMatrix4", GroupUniformMatrix4); EXPORT("groupUniformsSetFrame", GroupUniformsSetFrame); EXPORT("groupFbAcquire", GroupFbAcquire);
    EXPORT("groupFbRelease", GroupFbRelease); EXPORT("groupRenderSample", GroupRenderSample); EXPORT("groupPresent", GroupPresent);
    EXPORT("setGatherTarget", SetGatherTarget); EXPORT("fbScatterRows", FbScatterRows); EXPORT("displayPlanes", DisplayPlanes);
    return exports;
}
NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
