"""Python host mirror of the reference's CoreAPI_Base (lib/RenderSystem/core_api_base.h:81-119):
same method names, argument meaning and call-order contract, forwarding to the C ABI.
Errors surface as CoreError (the reference cores call FatalError and exit)."""
import ctypes
import numpy as np

from . import abi
from .capi import load_library

Converge, Restart = 0, 1


class CoreError(RuntimeError):
    pass


def _ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _arr(a, dtype):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


class RenderCore:
    """One render core on one CUDA device (CoreAPI_Base::CreateCoreAPI + Init)."""

    def __init__(self, device=-1):
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        self._check(self._lib.lh2b_create(ctypes.byref(self._h), device))
        self.width = self.height = self.spp = 0

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise CoreError(self._lib.lh2b_last_error().decode())

    def Shutdown(self):
        if self._h:
            self._lib.lh2b_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.Shutdown()
        except Exception:
            pass

    # -- CoreAPI_Base -----------------------------------------------------------------------
    def SetTarget(self, width, height, spp=1):
        self._check(self._lib.lh2b_set_target(self._h, width, height, spp))
        self.width, self.height, self.spp = width, height, spp

    def Setting(self, name, value):
        self._check(self._lib.lh2b_setting(self._h, name.encode(), float(value)))

    def SetProbePos(self, x, y):
        self._check(self._lib.lh2b_set_probe_pos(self._h, x, y))

    def SetTextures(self, texels_list):
        """texels_list: list of (array, storage) with array uint8[h,w,4] (ARGB32/NRM32) or float32[h,w,4] (ARGB128);
        each may already contain its MIP chain appended (reference layout). Returns the CoreTexDesc array."""
        n = len(texels_list)
        descs = np.zeros(max(n, 1), dtype=abi.CoreTexDesc)
        self._tex_keepalive = []
        for i, (tex, storage, w, h, mips, *flags) in enumerate(texels_list):
            tex = np.ascontiguousarray(tex)
            self._tex_keepalive.append(tex)
            descs[i]["data"] = tex.ctypes.data
            descs[i]["width"], descs[i]["height"] = w, h
            descs[i]["pixelCount"] = tex.size // 4
            descs[i]["MIPlevels"] = mips
            descs[i]["storage"] = storage
            descs[i]["flags"] = flags[0] if flags else 0
        self._check(self._lib.lh2b_set_textures(self._h, _ptr(descs), n))
        return descs[:n]

    def SetMaterials(self, materials):
        m = _arr(materials, abi.CoreMaterial)
        self._check(self._lib.lh2b_set_materials(self._h, _ptr(m), len(m)))

    def SetLights(self, tri=None, point=None, spot=None, directional=None):
        t, p = _arr(tri, abi.CoreLightTri), _arr(point, abi.CorePointLight)
        s, d = _arr(spot, abi.CoreSpotLight), _arr(directional, abi.CoreDirectionalLight)
        n = [0 if x is None else len(x) for x in (t, p, s, d)]
        self._check(self._lib.lh2b_set_lights(self._h, _ptr(t), n[0], _ptr(p), n[1], _ptr(s), n[2], _ptr(d), n[3]))

    def SetSkyData(self, pixels, width, height, world_to_light=None):
        px = _arr(pixels, np.float32)
        m = _arr(np.eye(4) if world_to_light is None else world_to_light, np.float32)
        self._check(self._lib.lh2b_set_sky(self._h, _ptr(px), width, height, _ptr(m)))

    def SetGeometry(self, mesh_idx, vertex_data, triangles=None):
        v = _arr(vertex_data, np.float32).reshape(-1, 4)
        tri_count = v.shape[0] // 3
        t = None if triangles is None else _arr(triangles, abi.CoreTri)
        self._check(self._lib.lh2b_set_geometry(self._h, mesh_idx, _ptr(v), v.shape[0], tri_count, _ptr(t)))

    def SetGeometryDevice(self, mesh_idx, d_vertex_ptr, tri_count, d_triangles_ptr=None):
        """SetGeometry from device pointers (ints): float4[3 * tri_count] and optionally CoreTri[tri_count]."""
        self._check(self._lib.lh2b_set_geometry_device(self._h, mesh_idx, ctypes.c_void_p(d_vertex_ptr), 3 * tri_count, tri_count,
                                                       ctypes.c_void_p(d_triangles_ptr) if d_triangles_ptr else None))

    def SetSkin(self, mesh_idx, joints4, weights4):
        j, w = _arr(joints4, np.uint32).reshape(-1, 4), _arr(weights4, np.float32).reshape(-1, 4)
        self._check(self._lib.lh2b_set_skin(self._h, mesh_idx, _ptr(j), _ptr(w), j.shape[0]))

    def SetPose(self, mesh_idx, joint_matrices):
        m = _arr(joint_matrices, np.float32).reshape(-1, 16)
        self._check(self._lib.lh2b_set_pose(self._h, mesh_idx, _ptr(m), m.shape[0]))

    def SetMorphTargets(self, mesh_idx, deltas4, normals4):
        d, n = _arr(deltas4, np.float32), _arr(normals4, np.float32)
        targets, verts = d.shape[0], d.shape[1]
        self._check(self._lib.lh2b_set_morph_targets(self._h, mesh_idx, _ptr(d), _ptr(n), targets, verts))

    def SetMorphWeights(self, mesh_idx, weights):
        w = _arr(weights, np.float32)
        self._check(self._lib.lh2b_set_morph_weights(self._h, mesh_idx, _ptr(w), w.shape[0]))

    def ReadGeometry(self, mesh_idx, tri_count):
        v, t = np.empty((3 * tri_count, 4), np.float32), np.empty(tri_count, abi.CoreTri)
        self._check(self._lib.lh2b_read_geometry(self._h, mesh_idx, _ptr(v), _ptr(t)))
        return v, t

    def SetInstance(self, instance_idx, mesh_idx, transform=None):
        m = _arr(np.eye(4) if transform is None else transform, np.float32)
        self._check(self._lib.lh2b_set_instance(self._h, instance_idx, mesh_idx, _ptr(m)))

    def FinalizeInstances(self):
        self._check(self._lib.lh2b_finalize_instances(self._h))

    def Render(self, view, converge=Restart, async_=False):
        v = _arr(view, abi.ViewPyramid)
        self._check(self._lib.lh2b_render(self._h, _ptr(v), int(converge), int(async_)))

    def WaitForRender(self):
        self._check(self._lib.lh2b_wait_for_render(self._h))

    def GetCoreStats(self):
        s = np.zeros(1, dtype=abi.CoreStats)
        self._check(self._lib.lh2b_get_stats(self._h, _ptr(s)))
        return s[0]

    # -- headless extras --------------------------------------------------------------------
    def ReadPixels(self, out=None):
        out = np.empty((self.height, self.width, 4), np.float32) if out is None else out
        self._check(self._lib.lh2b_read_pixels(self._h, _ptr(out)))
        return out

    def ReadPixelsAsync(self, pinned_out):
        """Enqueue the read-back of the last rendered frame into a page-locked array; overlaps the next Render."""
        self._check(self._lib.lh2b_read_pixels_async(self._h, _ptr(pinned_out)))

    def WaitReadPixels(self):
        self._check(self._lib.lh2b_wait_read_pixels(self._h))

    def ReadAccumulator(self):
        out = np.empty((self.height, self.width, 4), np.float32)
        self._check(self._lib.lh2b_read_accumulator(self._h, _ptr(out)))
        return out

    def AccumulatorDevicePtr(self):
        p, n = ctypes.c_void_p(), ctypes.c_int()
        self._check(self._lib.lh2b_accumulator_device_ptr(self._h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def SamplesTaken(self):
        """Samples accumulated per pixel so far, counted over the whole (possibly sharded) frame."""
        return self.AccumulatorDevicePtr()[1]

    def FinalizeExternal(self, accumulator, samples):
        """pixels = accumulator / samples for an externally reduced accumulator (torch CUDA tensor or raw device pointer)."""
        ptr = accumulator.data_ptr() if hasattr(accumulator, "data_ptr") else int(accumulator)
        self._check(self._lib.lh2b_finalize_external(self._h, ctypes.c_void_p(ptr), int(samples)))

    def SnapshotAccumulator(self, dst):
        """Enqueue a device-to-device copy of the accumulator behind the frame rendered last (torch CUDA tensor or device pointer)."""
        ptr = dst.data_ptr() if hasattr(dst, "data_ptr") else int(dst)
        self._check(self._lib.lh2b_snapshot_accumulator(self._h, ctypes.c_void_p(ptr)))

    def FinalizeExternalOn(self, accumulator, samples, pixels_out, stream):
        """Finalize kernel on the caller's CUDA stream (int handle): pixels_out = accumulator / samples, both device buffers."""
        a = accumulator.data_ptr() if hasattr(accumulator, "data_ptr") else int(accumulator)
        o = pixels_out.data_ptr() if hasattr(pixels_out, "data_ptr") else int(pixels_out)
        self._check(self._lib.lh2b_finalize_external_on(self._h, ctypes.c_void_p(a), int(samples), ctypes.c_void_p(o), ctypes.c_void_p(int(stream))))

    # -- multi-GPU frame gather over peer memory (csrc/gather.cu) ---------------------------------
    def GatherCreate(self, rank, world):
        g = ctypes.c_void_p()
        self._check(self._lib.lh2b_gather_create(self._h, rank, world, ctypes.byref(g)))
        return g

    def GatherExport(self, g):
        buf = ctypes.create_string_buffer(self._lib.lh2b_gather_handle_bytes())
        self._check(self._lib.lh2b_gather_export(g, buf))
        return buf.raw

    def GatherImport(self, g, handles_of_all_ranks):
        self._check(self._lib.lh2b_gather_import(g, ctypes.c_char_p(handles_of_all_ranks)))

    def GatherFrame(self, g, samples_total, pinned_out=None):
        p = None if pinned_out is None else (ctypes.c_void_p(pinned_out.data_ptr()) if hasattr(pinned_out, "data_ptr") else _ptr(pinned_out))
        self._check(self._lib.lh2b_gather_frame(g, int(samples_total), p))

    def GatherWait(self, g):
        self._check(self._lib.lh2b_gather_wait(g))

    def GatherJoin(self, g, stream):
        self._check(self._lib.lh2b_gather_join(g, ctypes.c_void_p(int(stream))))

    def GatherDestroy(self, g):
        self._check(self._lib.lh2b_gather_destroy(g))

    # ---- tile (row-band) sharding of one frame (csrc/tile_gather.cu) ----
    def SetRowBand(self, y0, y1, step_tile_rows=1):
        """Rows [y0, y1) only; step_tile_rows > 1: every step-th 4-row tile row of that range (interleaved bands)."""
        self._check(self._lib.lh2b_set_row_band_strided(self._h, int(y0), int(y1), int(step_tile_rows)))

    def TileCreate(self, rank, world):
        g = ctypes.c_void_p()
        self._check(self._lib.lh2b_tile_create(self._h, rank, world, ctypes.byref(g)))
        return g

    def TileExport(self, g):
        buf = ctypes.create_string_buffer(self._lib.lh2b_tile_handle_bytes())
        self._check(self._lib.lh2b_tile_export(g, buf))
        return buf.raw

    def TileImport(self, g, handles_of_all_ranks):
        self._check(self._lib.lh2b_tile_import(g, ctypes.c_char_p(handles_of_all_ranks)))

    def TileFrame(self, g):
        self._check(self._lib.lh2b_tile_frame(g))

    def TileWait(self, g):
        self._check(self._lib.lh2b_tile_wait(g))

    def TileRows(self, g):
        y0, y1, st = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._check(self._lib.lh2b_tile_rows(g, ctypes.byref(y0), ctypes.byref(y1), ctypes.byref(st)))
        return y0.value, y1.value, st.value

    def TileDestroy(self, g):
        self._check(self._lib.lh2b_tile_destroy(g))

    def SetSampleShard(self, first_sample, total_spp):
        self._check(self._lib.lh2b_set_sample_shard(self._h, first_sample, total_spp))

    def TraceRays(self, origins, directions):
        o, d = _arr(origins, np.float32).reshape(-1, 4), _arr(directions, np.float32).reshape(-1, 4)
        hits = np.empty((o.shape[0], 4), np.uint32)
        self._check(self._lib.lh2b_trace_rays(self._h, _ptr(o), _ptr(d), o.shape[0], _ptr(hits)))
        return hits

    def TraceShadowRays(self, origins, directions):
        o, d = _arr(origins, np.float32).reshape(-1, 4), _arr(directions, np.float32).reshape(-1, 4)
        occ = np.empty(o.shape[0], np.uint8)
        self._check(self._lib.lh2b_trace_shadow_rays(self._h, _ptr(o), _ptr(d), o.shape[0], _ptr(occ)))
        return occ

    def TraceRaysDevice(self, d_origins, d_directions, n, d_hits, repeat=1, timed=True):
        ms = ctypes.c_float(0)
        self._check(self._lib.lh2b_trace_rays_device(self._h, d_origins, d_directions, n, d_hits, repeat,
                                                     ctypes.byref(ms) if timed else None))
        return ms.value

    def TraceShadowRaysDevice(self, d_origins, d_directions, n, d_occ, repeat=1, timed=True):
        ms = ctypes.c_float(0)
        self._check(self._lib.lh2b_trace_shadow_rays_device(self._h, d_origins, d_directions, n, d_occ, repeat,
                                                            ctypes.byref(ms) if timed else None))
        return ms.value

    def ShadePaths(self, path_length, O4, D4, T4, hits, R0, shift, pass_, accumulator):
        """Parity hook (lh2b_shade_paths): the shade stage alone on host buffers. Returns (ext dict, shadow dict, accumulator)."""
        O4, D4, T4 = (_arr(a, np.float32).reshape(-1, 4) for a in (O4, D4, T4))
        hits = np.ascontiguousarray(hits).view(np.float32).reshape(-1, 4)
        n = O4.shape[0]
        out = [np.zeros((n, 4), np.float32) for _ in range(6)]
        acc = np.ascontiguousarray(accumulator, np.float32).copy()
        ne, ns = ctypes.c_int(0), ctypes.c_int(0)
        self._check(self._lib.lh2b_shade_paths(self._h, path_length, n, _ptr(O4), _ptr(D4), _ptr(T4), _ptr(hits), R0, shift, pass_,
                                               _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), ctypes.byref(ne),
                                               _ptr(out[3]), _ptr(out[4]), _ptr(out[5]), ctypes.byref(ns), _ptr(acc)))
        ext = dict(O=out[0][:ne.value], D=out[1][:ne.value], T=out[2][:ne.value])
        sh = dict(O=out[3][:ns.value], D=out[4][:ns.value], E=out[5][:ns.value])
        return ext, sh, acc

    def ShadePathsTime(self, path_length, O4, D4, T4, hits, R0, shift, pass_, runs=10):
        """(fastest, mean) ms of `runs` launches of the shade kernel on these path states (lh2b_shade_paths_time)."""
        o, d, t = (_arr(a, np.float32).reshape(-1, 4) for a in (O4, D4, T4))
        hb = np.ascontiguousarray(hits).view(np.float32).reshape(-1, 4)
        ms = np.zeros(2, np.float32)
        self._check(self._lib.lh2b_shade_paths_time(self._h, path_length, o.shape[0], _ptr(o), _ptr(d), _ptr(t), _ptr(hb), R0, shift, pass_, runs, _ptr(ms)))
        return float(ms[0]), float(ms[1])

    def ReadFilterBuffers(self):
        """Filter mode: (features uint32[h,w,4], worldPos, deltaDepth float32[h,w,4], accumulator float32[2,h,w,4]) of the last frame."""
        h, w = self.height, self.width
        f, wp, dd, acc = np.zeros((h, w, 4), np.uint32), np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32), np.zeros((2, h, w, 4), np.float32)
        self._check(self._lib.lh2b_read_filter_buffers(self._h, _ptr(f), _ptr(wp), _ptr(dd), _ptr(acc)))
        return f, wp, dd, acc

    def ReadFilterHistory(self):
        """Filter mode: dict of what the last frame's chain left behind (float32[h,w,4]; motion float32[h,w,2]) - moments, phase-1 output,
        TAA image, phase-3 output, motion vectors (for tests)."""
        h, w = self.height, self.width
        out = {k: np.zeros((h, w, 4), np.float32) for k in ("moments", "phase1", "taa", "phase3")}
        out["motion"] = np.zeros((h, w, 2), np.float32)
        self._check(self._lib.lh2b_read_filter_history(self._h, *[_ptr(out[k]) for k in ("moments", "phase1", "taa", "phase3", "motion")]))
        return out

    def DebugReadTable(self, name):
        """One of the core's device tables as raw bytes (uint8 array); see lh2b_debug_read_table."""
        n = ctypes.c_size_t()
        self._check(self._lib.lh2b_debug_read_table(self._h, name.encode(), None, 0, ctypes.byref(n)))
        out = np.zeros(n.value, np.uint8)
        if n.value:
            self._check(self._lib.lh2b_debug_read_table(self._h, name.encode(), _ptr(out), n.value, ctypes.byref(n)))
        return out

    def FilterChain(self, io):
        """Parity hook (lh2b_filter_chain): io is a ctypes structure laid out like lh2b_filter_io."""
        self._check(self._lib.lh2b_filter_chain(self._h, ctypes.byref(io)))

    def Stream(self):
        p = ctypes.c_void_p()
        self._check(self._lib.lh2b_stream(self._h, ctypes.byref(p)))
        return p.value

    def GetFrameStats(self):
        s = np.zeros(1, dtype=abi.FrameStats)
        self._check(self._lib.lh2b_get_frame_stats(self._h, _ptr(s)))
        return s[0]

    TRACE_STATS = ("rays", "nodeSteps", "triTests", "instanceEntries", "iterations", "nodePhases", "triPhases", "nodeLanes", "triLanes")

    def TraceStatsEnable(self, on=True):
        """Work counters of the traversal kernels (measurement): counting instantiations run while enabled."""
        self._check(self._lib.lh2b_trace_stats_enable(self._h, 1 if on else 0))

    def TraceStatsRead(self, reset=True):
        out = np.zeros(9, np.uint64)
        self._check(self._lib.lh2b_trace_stats_read(self._h, _ptr(out), 1 if reset else 0))
        return dict(zip(self.TRACE_STATS, (int(x) for x in out)))

    def GetBvhStats(self, mesh_idx=-1):
        s = np.zeros(1, dtype=abi.BvhStats)
        self._check(self._lib.lh2b_get_bvh_stats(self._h, mesh_idx, _ptr(s)))
        return s[0]
