// rmb_group.cpp -- one handle for the GPUs of one box (include/rmb.h, "device groups").
//
// SURVEY.md 8b asks for `ctx_create(device_ids[], n)`: one context that owns the streams, the program cache and
// the buffer pool of every device, so that a single-threaded host (the reference's TypeScript renderer,
// client/src/index.tsx:236-263 -> renderer/RenderJobExecutor.tsx:77-341) reaches all GPUs through one object.
// A group is n member contexts (rmb_ctx, member i renders the row tiles t with t % n == i) plus what
// ties them together in ONE process, with neither torch nor NCCL:
//   * every entry point of the single-device API once more at group level, fanned out to the members
//     (programs are compiled on one host thread per member, everything else is enqueued in a loop);
//   * the assembled frame lives on member 0's device; the other devices map it with cudaDeviceEnablePeerAccess and
//     their display kernels store straight into it over NVLink (rmb_ctx_set_gather_target, display.cu);
//   * completion is ordered on the device: an event per member, member 0's stream waits for all of them, the host
//     only waits for member 0's readback.
// Frames whose display pass blurs (full mode with depth of field, display.frag:25-55) scatter the colour and
// normal+dofRadius rows to member 0, which presents the assembled planes (rmb_fb_scatter_rows, rmb_display_planes).
// Everything here is a client of the public single-device ABI plus the CUDA runtime.
#include <cuda_runtime.h>

#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "../../include/rmb.h"

struct rmb_group_program {
    rmb_group* group = nullptr;
    std::vector<rmb_program*> member;
    int render_mode = 0;      // last value of the renderMode uniform set through the group (1 = preview)
};

struct rmb_group_fb {
    rmb_group* group = nullptr;
    int width = 0, height = 0;
    int64_t frameid = 0;
    std::vector<rmb_fb*> member;
    bool may_blur = false;    // some sample since the set was acquired was drawn in full mode (nd.w can be non-zero)
};

struct rmb_group {
    std::vector<int> device;
    std::vector<rmb_ctx*> ctx;
    std::vector<cudaEvent_t> done;        // member i has stored its rows of the frame being presented
    cudaEvent_t consumed = nullptr;       // member 0 has finished reading the assembled buffers
    bool consumed_valid = false;
    int tile_rows = 16;
    std::string last_error;
    // assembled full-frame buffers on member 0's device
    void* rgba8 = nullptr; size_t rgba8_bytes = 0;
    void* color = nullptr; size_t color_bytes = 0;
    void* nd = nullptr; size_t nd_bytes = 0;
    std::vector<void*> depth_stage;       // pinned host staging per member (depth readback)
    std::vector<size_t> depth_stage_bytes;
    std::vector<std::unique_ptr<rmb_group_program>> programs;
    std::map<std::tuple<int, int, int64_t>, std::unique_ptr<rmb_group_fb>> fbs;
    // released sets stay addressable like the members' "purgatory" (LoadRenderJobContext.tsx:176, cap 3): doRenderJob
    // deletes its framebuffers BEFORE the final present (RenderJobExecutor.tsx:333-338)
    std::vector<std::unique_ptr<rmb_group_fb>> released;
};

namespace {

thread_local std::string g_group_error;

rmb_status gfail(rmb_group* g, rmb_status st, const std::string& msg) {
    if (g) g->last_error = msg; else g_group_error = msg;
    return st;
}
rmb_status member_fail(rmb_group* g, int i, rmb_status st) {
    const char* e = rmb_last_error(g->ctx[i]);
    return gfail(g, st, "member " + std::to_string(i) + " (device " + std::to_string(g->device[i]) + "): " + (e ? e : "error"));
}
#define G_CUDA(g, expr)                                                                                         \
    do {                                                                                                        \
        cudaError_t e__ = (expr);                                                                               \
        if (e__ != cudaSuccess) return gfail(g, RMB_ERR_GENERAL, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

cudaStream_t stream_of(rmb_group* g, int i) { return (cudaStream_t)rmb_ctx_stream(g->ctx[i]); }

rmb_status ensure(rmb_group* g, void** p, size_t* have, size_t want) {
    if (*have >= want) return RMB_OK;
    // the old buffer may still be read or written by queued work of any member
    for (size_t i = 0; i < g->ctx.size(); i++) rmb_sync(g->ctx[i]);
    if (*p) rmb_device_free(g->ctx[0], *p);
    *p = rmb_device_alloc(g->ctx[0], want);
    *have = *p ? want : 0;
    if (!*p) return gfail(g, RMB_ERR_GENERAL, "rmb_device_alloc failed for the assembled frame");
    return RMB_OK;
}

// every member but 0 waits until member 0 has consumed the assembled buffers of the previous present
rmb_status wait_consumed(rmb_group* g) {
    if (!g->consumed_valid) return RMB_OK;
    for (size_t i = 1; i < g->ctx.size(); i++) {
        G_CUDA(g, cudaSetDevice(g->device[i]));
        G_CUDA(g, cudaStreamWaitEvent(stream_of(g, (int)i), g->consumed, 0));
    }
    return RMB_OK;
}

// member i records "my rows are stored"; member 0's stream waits for every member
rmb_status join_on_member0(rmb_group* g) {
    for (size_t i = 1; i < g->ctx.size(); i++) {
        G_CUDA(g, cudaSetDevice(g->device[i]));
        G_CUDA(g, cudaEventRecord(g->done[i], stream_of(g, (int)i)));
    }
    G_CUDA(g, cudaSetDevice(g->device[0]));
    for (size_t i = 1; i < g->ctx.size(); i++) G_CUDA(g, cudaStreamWaitEvent(stream_of(g, 0), g->done[i], 0));
    return RMB_OK;
}

}  // namespace

extern "C" {

const char* rmb_group_last_error(rmb_group* g) { return g ? g->last_error.c_str() : g_group_error.c_str(); }

rmb_group* rmb_group_create(const int* devices, int n, int tile_rows) {
    if (!devices || n < 1 || n > 64 || tile_rows < 1) { gfail(nullptr, RMB_ERR_INVALID, "rmb_group_create: bad arguments"); return nullptr; }
    std::unique_ptr<rmb_group> g(new rmb_group);
    g->tile_rows = tile_rows;
    auto destroy_partial = [&]() {
        for (rmb_ctx* c : g->ctx) rmb_ctx_destroy(c);
        for (cudaEvent_t e : g->done) if (e) cudaEventDestroy(e);
        if (g->consumed) cudaEventDestroy(g->consumed);
    };
    for (int i = 0; i < n; i++) {
        rmb_ctx* c = rmb_ctx_create(devices[i], i, n, tile_rows);
        if (!c) {
            const char* e = rmb_last_error(nullptr);
            gfail(nullptr, RMB_ERR_GENERAL, std::string("rmb_group_create: device ") + std::to_string(devices[i]) + ": " + (e ? e : "no context"));
            destroy_partial();
            return nullptr;
        }
        g->device.push_back(devices[i]);
        g->ctx.push_back(c);
        g->done.push_back(nullptr);
    }
    for (int i = 0; i < n; i++) {
        cudaSetDevice(devices[i]);
        if (cudaEventCreateWithFlags(&g->done[i], cudaEventDisableTiming) != cudaSuccess) {
            gfail(nullptr, RMB_ERR_GENERAL, "rmb_group_create: cudaEventCreate failed");
            destroy_partial();
            return nullptr;
        }
        if (devices[i] == devices[0]) continue;
        // stores into member 0's memory: peer access from every other device (NVLink / NVSwitch on one box)
        int can = 0;
        cudaDeviceCanAccessPeer(&can, devices[i], devices[0]);
        cudaError_t e = can ? cudaDeviceEnablePeerAccess(devices[0], 0) : cudaErrorPeerAccessUnsupported;
        if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
        if (e != cudaSuccess) {
            gfail(nullptr, RMB_ERR_GENERAL, "rmb_group_create: device " + std::to_string(devices[i]) + " cannot access device " +
                                                std::to_string(devices[0]) + " (" + cudaGetErrorString(e) + ")");
            cudaGetLastError();
            destroy_partial();
            return nullptr;
        }
    }
    cudaSetDevice(devices[0]);
    if (cudaEventCreateWithFlags(&g->consumed, cudaEventDisableTiming) != cudaSuccess) {
        gfail(nullptr, RMB_ERR_GENERAL, "rmb_group_create: cudaEventCreate failed");
        destroy_partial();
        return nullptr;
    }
    g->depth_stage.assign(n, nullptr);
    g->depth_stage_bytes.assign(n, 0);
    return g.release();
}

void rmb_group_destroy(rmb_group* g) {
    if (!g) return;
    for (rmb_ctx* c : g->ctx) rmb_sync(c);
    for (rmb_ctx* c : g->ctx) rmb_ctx_set_gather_target(c, nullptr, 0);
    if (g->rgba8) rmb_device_free(g->ctx[0], g->rgba8);
    if (g->color) rmb_device_free(g->ctx[0], g->color);
    if (g->nd) rmb_device_free(g->ctx[0], g->nd);
    for (void* p : g->depth_stage) if (p) rmb_host_free(p);
    for (size_t i = 0; i < g->done.size(); i++) { cudaSetDevice(g->device[i]); if (g->done[i]) cudaEventDestroy(g->done[i]); }
    cudaSetDevice(g->device[0]);
    if (g->consumed) cudaEventDestroy(g->consumed);
    for (rmb_ctx* c : g->ctx) rmb_ctx_destroy(c);
    delete g;
}

int rmb_group_size(rmb_group* g) { return g ? (int)g->ctx.size() : 0; }
rmb_ctx* rmb_group_ctx(rmb_group* g, int member) { return (g && member >= 0 && member < (int)g->ctx.size()) ? g->ctx[member] : nullptr; }

rmb_status rmb_group_sync(rmb_group* g) {
    if (!g) return RMB_ERR_INVALID;
    for (size_t i = 0; i < g->ctx.size(); i++) {
        rmb_status st = rmb_sync(g->ctx[i]);
        if (st != RMB_OK) return member_fail(g, (int)i, st);
    }
    return RMB_OK;
}

rmb_status rmb_group_program_get(rmb_group* g, const char* scene_glsl, size_t scene_len, int flavour, const rmb_spec_uniform* spec,
                                 int n_spec, rmb_group_program** out_program, char* err_type, char* infolog, size_t infolog_cap) {
    if (out_program) *out_program = nullptr;
    if (!g || !scene_glsl || !out_program) return gfail(g, RMB_ERR_INVALID, "rmb_group_program_get: null argument");
    const int n = (int)g->ctx.size();
    std::vector<rmb_program*> prog(n, nullptr);
    std::vector<rmb_status> st(n, RMB_OK);
    std::vector<std::string> etype(n, std::string(16, '\0')), log(n, std::string(1 << 16, '\0'));
    // lowering + NVRTC + module load take seconds per device and touch only the member's own state: one host thread each
    std::vector<std::thread> workers;
    for (int i = 0; i < n; i++)
        workers.emplace_back([&, i]() {
            st[i] = rmb_program_get(g->ctx[i], scene_glsl, scene_len, flavour, spec, n_spec, &prog[i], &etype[i][0], &log[i][0], log[i].size());
        });
    for (auto& w : workers) w.join();
    for (int i = 0; i < n; i++)
        if (st[i] != RMB_OK) {
            // compile errors are values (ShaderCache.tsx:8-11): hand out the first member's diagnosis
            if (err_type) { strncpy(err_type, etype[i].c_str(), 15); err_type[15] = 0; }
            if (infolog && infolog_cap) { strncpy(infolog, log[i].c_str(), infolog_cap - 1); infolog[infolog_cap - 1] = 0; }
            g->last_error = log[i].c_str();
            return st[i];
        }
    if (err_type) err_type[0] = 0;
    if (infolog && infolog_cap) infolog[0] = 0;
    for (auto& p : g->programs)
        if (p->member == prog) { *out_program = p.get(); return RMB_OK; }     // cache hit on every member
    std::unique_ptr<rmb_group_program> gp(new rmb_group_program);
    gp->group = g;
    gp->member = prog;
    *out_program = gp.get();
    g->programs.push_back(std::move(gp));
    return RMB_OK;
}

rmb_program* rmb_group_program_member(rmb_group_program* p, int member) {
    return (p && member >= 0 && member < (int)p->member.size()) ? p->member[member] : nullptr;
}

static void note_render_mode(rmb_group_program* p, const char* name, int type, const void* data) {
    if (strcmp(name, "renderMode") != 0) return;
    if (type == RMB_UNIFORM_F) p->render_mode = (int)*(const float*)data;
    else p->render_mode = *(const int32_t*)data;
}

rmb_status rmb_group_uniform_set(rmb_group_program* p, const char* name, int type, int count, const void* data) {
    if (!p || !name || !data) return RMB_ERR_INVALID;
    for (size_t i = 0; i < p->member.size(); i++) {
        rmb_status st = rmb_uniform_set(p->member[i], name, type, count, data);
        if (st != RMB_OK) return member_fail(p->group, (int)i, st);
    }
    note_render_mode(p, name, type, data);
    return RMB_OK;
}

rmb_status rmb_group_uniform_set_array(rmb_group_program* p, const char* name, int type, int components, int n_elements, const void* data) {
    if (!p || !name || !data) return RMB_ERR_INVALID;
    for (size_t i = 0; i < p->member.size(); i++) {
        rmb_status st = rmb_uniform_set_array(p->member[i], name, type, components, n_elements, data);
        if (st != RMB_OK) return member_fail(p->group, (int)i, st);
    }
    return RMB_OK;
}

rmb_status rmb_group_uniform_matrix4(rmb_group_program* p, const char* name, const float* m16_column_major) {
    if (!p || !name || !m16_column_major) return RMB_ERR_INVALID;
    for (size_t i = 0; i < p->member.size(); i++) {
        rmb_status st = rmb_uniform_matrix4(p->member[i], name, m16_column_major);
        if (st != RMB_OK) return member_fail(p->group, (int)i, st);
    }
    return RMB_OK;
}

rmb_status rmb_group_uniforms_set_frame(rmb_group_program* p, const rmb_frame_uniforms* u) {
    if (!p || !u) return RMB_ERR_INVALID;
    for (size_t i = 0; i < p->member.size(); i++) {
        rmb_status st = rmb_uniforms_set_frame(p->member[i], u);
        if (st != RMB_OK) return member_fail(p->group, (int)i, st);
    }
    p->render_mode = u->renderMode;
    return RMB_OK;
}

rmb_group_fb* rmb_group_fb_acquire(rmb_group* g, int width, int height, int64_t frameid) {
    if (!g || width < 1 || height < 1) { gfail(g, RMB_ERR_INVALID, "rmb_group_fb_acquire: bad size"); return nullptr; }
    const auto key = std::make_tuple(width, height, frameid);
    auto hit = g->fbs.find(key);
    if (hit != g->fbs.end()) return hit->second.get();
    std::unique_ptr<rmb_group_fb> fb;
    for (size_t k = 0; k < g->released.size(); k++)
        if (g->released[k]->width == width && g->released[k]->height == height && g->released[k]->frameid == frameid) {
            fb = std::move(g->released[k]);          // same frameid: contents (and the blur flag) kept
            g->released.erase(g->released.begin() + k);
            fb->member.clear();
            break;
        }
    if (!fb) {
        fb.reset(new rmb_group_fb);
        fb->group = g; fb->width = width; fb->height = height; fb->frameid = frameid;
    }
    for (size_t i = 0; i < g->ctx.size(); i++) {
        rmb_fb* m = rmb_fb_acquire(g->ctx[i], width, height, frameid);
        if (!m) {
            member_fail(g, (int)i, RMB_ERR_GENERAL);
            for (size_t k = 0; k < i; k++) rmb_fb_release(g->ctx[k], width, height, frameid);
            return nullptr;
        }
        fb->member.push_back(m);
    }
    rmb_group_fb* raw = fb.get();
    g->fbs[key] = std::move(fb);
    return raw;
}

void rmb_group_fb_release(rmb_group* g, int width, int height, int64_t frameid) {
    if (!g) return;
    auto hit = g->fbs.find(std::make_tuple(width, height, frameid));
    if (hit == g->fbs.end()) return;
    for (size_t i = 0; i < g->ctx.size(); i++) rmb_fb_release(g->ctx[i], width, height, frameid);
    g->released.push_back(std::move(hit->second));
    g->fbs.erase(hit);
    if (g->released.size() > 3) g->released.erase(g->released.begin());
}

rmb_fb* rmb_group_fb_member(rmb_group_fb* fb, int member) {
    return (fb && member >= 0 && member < (int)fb->member.size()) ? fb->member[member] : nullptr;
}

rmb_status rmb_group_render_sample(rmb_group* g, rmb_group_program* p, rmb_group_fb* fb, int sx, int sy, int sw, int sh) {
    if (!g || !p || !fb || p->group != g || fb->group != g) return gfail(g, RMB_ERR_INVALID, "rmb_group_render_sample: bad handle");
    if (p->render_mode != 1) fb->may_blur = true;
    for (size_t i = 0; i < g->ctx.size(); i++) {
        rmb_status st = rmb_render_sample(g->ctx[i], p->member[i], fb->member[i], sx, sy, sw, sh);
        if (st != RMB_OK) return member_fail(g, (int)i, st);
    }
    return RMB_OK;
}

rmb_status rmb_group_present_device(rmb_group* g, rmb_group_fb* fb, float brightness, void** rgba8_device) {
    if (!g || !fb || fb->group != g) return gfail(g, RMB_ERR_INVALID, "rmb_group_present_device: bad handle");
    const int n = (int)g->ctx.size();
    const size_t px = (size_t)fb->width * (size_t)fb->height;
    rmb_status st = ensure(g, &g->rgba8, &g->rgba8_bytes, px * 4);
    if (st != RMB_OK) return st;
    if ((st = wait_consumed(g)) != RMB_OK) return st;
    if (!fb->may_blur || n == 1) {
        // the display kernel of every member stores its rows a second time at their global row of the assembled frame
        for (int i = 0; i < n; i++) {
            rmb_ctx_set_gather_target(g->ctx[i], g->rgba8, g->rgba8_bytes);
            st = rmb_present_device(g->ctx[i], fb->member[i], brightness);
            rmb_ctx_set_gather_target(g->ctx[i], nullptr, 0);
            if (st != RMB_OK) return member_fail(g, i, st);
        }
        if ((st = join_on_member0(g)) != RMB_OK) return st;
    } else {
        // the blur reads up to 16 rows either side (REPEAT-wrapped): assemble the accumulators, present on member 0
        if ((st = ensure(g, &g->color, &g->color_bytes, px * 16)) != RMB_OK) return st;
        if ((st = ensure(g, &g->nd, &g->nd_bytes, px * 8)) != RMB_OK) return st;
        for (int i = 0; i < n; i++) {
            if ((st = rmb_fb_scatter_rows(g->ctx[i], fb->member[i], 0, g->color)) != RMB_OK) return member_fail(g, i, st);
            if ((st = rmb_fb_scatter_rows(g->ctx[i], fb->member[i], 1, g->nd)) != RMB_OK) return member_fail(g, i, st);
        }
        if ((st = join_on_member0(g)) != RMB_OK) return st;
        if ((st = rmb_display_planes(g->ctx[0], g->color, g->nd, g->rgba8, fb->width, fb->height, brightness)) != RMB_OK) return member_fail(g, 0, st);
    }
    G_CUDA(g, cudaSetDevice(g->device[0]));
    G_CUDA(g, cudaEventRecord(g->consumed, stream_of(g, 0)));
    g->consumed_valid = true;
    if (rgba8_device) *rgba8_device = g->rgba8;
    return RMB_OK;
}

rmb_status rmb_group_present(rmb_group* g, rmb_group_fb* fb, float brightness, uint8_t* rgba8_host, float* depth_host) {
    rmb_status st = rmb_group_present_device(g, fb, brightness, nullptr);
    if (st != RMB_OK) return st;
    const int n = (int)g->ctx.size();
    const size_t px = (size_t)fb->width * (size_t)fb->height;
    G_CUDA(g, cudaSetDevice(g->device[0]));
    if (rgba8_host) {
        G_CUDA(g, cudaMemcpyAsync(rgba8_host, g->rgba8, px * 4, cudaMemcpyDeviceToHost, stream_of(g, 0)));
        G_CUDA(g, cudaEventRecord(g->consumed, stream_of(g, 0)));
    }
    if (depth_host) {
        // hit depth of the latest sample (fp32): every member's rows through pinned staging, placed at
        // their global rows by the host
        for (int i = 0; i < n; i++) {
            const size_t bytes = rmb_fb_plane_bytes(fb->member[i], 3);
            if (!bytes) continue;
            if (g->depth_stage_bytes[i] < bytes) {
                if (g->depth_stage[i]) rmb_host_free(g->depth_stage[i]);
                g->depth_stage[i] = rmb_host_alloc(bytes);
                g->depth_stage_bytes[i] = g->depth_stage[i] ? bytes : 0;
                if (!g->depth_stage[i]) return gfail(g, RMB_ERR_GENERAL, "rmb_host_alloc failed");
            }
            G_CUDA(g, cudaSetDevice(g->device[i]));
            G_CUDA(g, cudaMemcpyAsync(g->depth_stage[i], rmb_fb_device_ptr(fb->member[i], 3), bytes, cudaMemcpyDeviceToHost, stream_of(g, i)));
        }
        for (int i = 0; i < n; i++) {
            if ((st = rmb_sync(g->ctx[i])) != RMB_OK) return member_fail(g, i, st);
            const int rows = rmb_fb_local_rows(fb->member[i]);
            const float* src = (const float*)g->depth_stage[i];
            for (int r = 0; r < rows; r++)
                memcpy(depth_host + (size_t)rmb_fb_global_row(fb->member[i], r) * fb->width, src + (size_t)r * fb->width, (size_t)fb->width * 4);
        }
    }
    if ((st = rmb_sync(g->ctx[0])) != RMB_OK) return member_fail(g, 0, st);
    return RMB_OK;
}

}  // extern "C"
