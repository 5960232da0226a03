// RenderJobExecutorB200.ts -- the reference's renderer/RenderJobExecutor.tsx with the WebGL2 calls
// re-pointed at the rmb N-API addon (integration/rmb_napi.c -> include/rmb.h).  Same exported names,
// same generator protocol, same error values; RenderJobSchema.tsx is used unchanged.
// NOT COMPILED IN THIS IMAGE (no tsc/Node); raymarching_engine_b200/executor.py is the tested mirror.
import { halton } from "../util/Halton";                       // unchanged reference module
import { RenderJobSchema } from "./RenderJobSchema";           // unchanged reference module
import { UniformData, u } from "./Uniforms";                   // types + constructors only
// eslint-disable-next-line @typescript-eslint/no-var-requires
const rmb = require("../../native/rmb.node");

export type ShaderError = { type: "vertex" | "fragment" | "program" | "general"; infoLog: string };
export type RenderJobFramebufferInfo = { handle: unknown; width: number; height: number; frameid: number; localRows: number };
// what takes the place of the WebGL2RenderingContext in the `gl` position of every callback: the native handle
// (a single-device context or a device group) - opaque to the host, exactly like `gl` is to present()'s callers
export type B200Device = { handle: unknown; group: boolean };
export type RenderJobContext = {
  gl: B200Device;                      // RenderJobExecutor.tsx:33 `gl: WebGL2RenderingContext`
  handle: unknown;
  flavour: 0 | 1;
  programCache: { getProgram: (scene: string, spec: Record<string, UniformData>) => { program: unknown } | ShaderError };
  fbo: {
    create: (width: number, height: number, frameid: number) => RenderJobFramebufferInfo | undefined;
    delete: (width: number, height: number, frameid: number) => void;
  };
};

const TYPE = { f: 0, i: 1, ui: 2 } as const;

// loadRenderJobContext(gl) -> loadRenderJobContext(device)        LoadRenderJobContext.tsx:268-287
export function loadRenderJobContext(device = 0, rank = 0, nRanks = 1, tileRows = 16, flavour: 0 | 1 = 0): RenderJobContext | undefined {
  const handle = rmb.ctxCreate(device, rank, nRanks, tileRows);
  if (handle === undefined) return undefined;
  return {
    gl: { handle, group: false },
    handle,
    flavour,
    programCache: {
      getProgram: (scene, spec) =>
        rmb.programGet(handle, scene, flavour, Object.entries(spec).map(([name, d]) => ({ name, type: TYPE[d.type], data: d.data }))),
    },
    fbo: {
      create: (width, height, frameid) => {
        const h = rmb.fbAcquire(handle, width, height, frameid);
        return h === undefined ? undefined : { handle: h, width, height, frameid, localRows: rmb.fbLocalRows(h) };
      },
      delete: (width, height, frameid) => rmb.fbRelease(handle, width, height, frameid),
    },
  };
}

// loadRenderJobContext over every GPU of the box (include/rmb.h device groups): the same RenderJobContext shape, so
// doRenderJob / makePresenter below drive 1 or 8 GPUs unchanged; present() receives the assembled full frame.
export function loadRenderJobGroup(devices: number[], tileRows = 16, flavour: 0 | 1 = 0): RenderJobContext | undefined {
  const handle = rmb.groupCreate(devices, tileRows);
  if (handle === undefined) return undefined;
  return {
    gl: { handle, group: true },
    handle,
    flavour,
    programCache: {
      getProgram: (scene, spec) =>
        rmb.groupProgramGet(handle, scene, flavour, Object.entries(spec).map(([name, d]) => ({ name, type: TYPE[d.type], data: d.data }))),
    },
    fbo: {
      create: (width, height, frameid) => {
        const h = rmb.groupFbAcquire(handle, width, height, frameid);
        return h === undefined ? undefined : { handle: h, width, height, frameid, localRows: height };
      },
      delete: (width, height, frameid) => rmb.groupFbRelease(handle, width, height, frameid),
    },
  };
}

const renderJobHalton2 = halton(2);                            // RenderJobExecutor.tsx:70-71
const renderJobHalton3 = halton(3);
const genErr = (infoLog: string): ShaderError => ({ type: "general", infoLog });

function setUniforms(gl: B200Device, program: unknown, uniforms: Record<string, UniformData>) {   // Uniforms.tsx:34-46
  for (const [name, d] of Object.entries(uniforms)) {
    const arr = d.type === "f" ? new Float32Array(d.data) : d.type === "i" ? new Int32Array(d.data) : new Uint32Array(d.data);
    (gl.group ? rmb.groupUniformSet : rmb.uniformSet)(program, name, TYPE[d.type], d.count, arr);
  }
}

export async function doRenderJob(schema: RenderJobSchema, context: RenderJobContext) {
  const framebuffers = context.fbo.create(schema.render.width, schema.render.height, schema.render.frameid);
  if (!framebuffers) return function* () { return { success: false, why: genErr("Failed to load framebuffers.") }; };
  const got = context.programCache.getProgram(schema.sdfShaderSource, schema.customShaderParameters);
  if (!("program" in got)) return function* () { return { success: false, why: got }; };
  const program = got.program;
  let samplesRenderedSoFar = 0;
  const gl = context.gl;
  const uniformSetArray = gl.group ? rmb.groupUniformSetArray : rmb.uniformSetArray;
  // the callback keeps the reference's five-argument shape (RenderJobExecutor.tsx:139-147), `gl` first
  return function* (present: (gl: B200Device, schema: RenderJobSchema, context: RenderJobContext, framebuffers: RenderJobFramebufferInfo, samplesSoFar: number) => void) {
    const r = schema.render;
    for (let yPartitions = 0; yPartitions < r.subdivisions; yPartitions++) {
      for (let xPartitions = 0; xPartitions < r.subdivisions; xPartitions++) {
        for (let sampleIndex = 0; sampleIndex < r.samplesPerPixel; sampleIndex++) {
          if (samplesRenderedSoFar % r.sampleYieldInterval == 0) { present(gl, schema, context, framebuffers, samplesRenderedSoFar); yield; }
          const x1 = Math.floor((r.width / r.subdivisions) * xPartitions), y1 = Math.floor((r.height / r.subdivisions) * yPartitions);
          const x2 = Math.ceil((r.width / r.subdivisions) * (xPartitions + 1)), y2 = Math.ceil((r.height / r.subdivisions) * (yPartitions + 1));
          const counts = schema.reflectionIterationCounts;
          const mode = schema.camera.mode;
          setUniforms(gl, program, {                                          // RenderJobExecutor.tsx:212-264
            blendWithPreviousFactor: u.float(r.blendWithPreviousFrameFactor),
            randNoise: u.vec2(renderJobHalton2.next().value, renderJobHalton3.next().value),
            position: u.vec3(...schema.camera.position),
            dofAmount: u.float(schema.dof.amount), dofFocalPlaneDistance: u.float(schema.dof.distance),
            cameraMode: u.int(["perspective", "orthographic", "panoramic"].indexOf(mode.type)),
            fov: u.float(mode.type == "perspective" ? mode.fov : mode.type == "orthographic" ? mode.size : 1),
            reflections: u.float(counts.length), aspect: u.float(r.width / r.height), fogDensity: u.float(schema.fogDensity),
            exposure: u.float(r.exposure / r.samplesPerPixel), blendMode: u.int(r.blendMode == "additive" ? 1 : 0),
            renderMode: u.int(r.renderMode == "preview" ? 1 : 0), lightCount: u.int(schema.lights.length),
            showDofFocalPlane: u.int(schema.dof.showFocusedArea ? 1 : 0),
          });
          setUniforms(gl, program, schema.customShaderParameters);            // :266
          uniformSetArray(program, "raymarchingStepCountsArray", 0, 1, counts.length, new Float32Array(counts));   // :268-274
          if (schema.lights.length > 0) {                                     // :276-291
            const L = schema.lights;
            uniformSetArray(program, "lightPositions", 0, 3, L.length, new Float32Array(L.flatMap((l) => (l.type == "point" ? l.position : l.direction))));
            uniformSetArray(program, "lightColors", 0, 3, L.length, new Float32Array(L.flatMap((l) => l.color)));
            uniformSetArray(program, "lightSizes", 0, 1, L.length, new Float32Array(L.map((l) => (l.type == "point" ? l.size : 0))));
          }
          (gl.group ? rmb.groupUniformMatrix4 : rmb.uniformMatrix4)(program, "rotation", new Float32Array(schema.camera.rotation as unknown as number[]));   // :293-297
          // gl.scissor(x1, y1, x2, y2) + drawArrays + blit                    :181-326
          if ((gl.group ? rmb.groupRenderSample : rmb.renderSample)(context.handle, program, framebuffers.handle, x1, y1, x2, y2) != 0)
            return { success: false, why: genErr((gl.group ? rmb.groupLastError : rmb.lastError)(context.handle)) };
          samplesRenderedSoFar++;
        }
      }
    }
    context.fbo.delete(r.width, r.height, r.frameid);                         // :333-337
    present(gl, schema, context, framebuffers, samplesRenderedSoFar);         // :338
    return { success: true };
  };
}

// makePresenter (index.tsx:25-59): the canvas becomes a pair of caller-owned typed arrays.
export function makePresenter(samplesUpToThisPoint: number, rgba8: Uint8Array, depth: Float32Array | null) {
  return function present(gl: B200Device, _schema: RenderJobSchema, _context: RenderJobContext, framebuffers: RenderJobFramebufferInfo, _samplesSoFar: number) {
    (gl.group ? rmb.groupPresent : rmb.present)(gl.handle, framebuffers.handle, 1 / samplesUpToThisPoint, rgba8, depth);
  };
}
