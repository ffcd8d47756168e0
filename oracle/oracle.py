"""TEST INFRASTRUCTURE ONLY: ctypes/numpy front end of oracle/liboracle.so (see ngp_oracle.h).

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs only. Builds the library on first use with `make -C oracle liboracle.so`.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MLP_PARAMS = 10240
GRID_CELLS = 128 ** 3


class Pcg32(C.Structure):
    _fields_ = [("state", C.c_uint64), ("inc", C.c_uint64)]


class Model(C.Structure):
    _fields_ = [("n_levels", C.c_uint32), ("log2_hashmap_size", C.c_uint32), ("base_resolution", C.c_uint32),
                ("per_level_scale", C.c_float), ("offsets", C.c_uint32 * 33), ("scales", C.c_float * 32),
                ("n_grid_params", C.c_uint32)]


class Image(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("w", C.c_int32), ("h", C.c_int32), ("fx", C.c_float), ("fy", C.c_float),
                ("cx", C.c_float), ("cy", C.c_float), ("xform", C.c_float * 12), ("lens_mode", C.c_int32), ("lens_params", C.c_float * 7), ("image_type", C.c_int32)]


class Optimizer(C.Structure):
    _fields_ = [("learning_rate", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("epsilon", C.c_float),
                ("l2_reg", C.c_float), ("ema_decay", C.c_float), ("decay_start", C.c_uint32), ("decay_interval", C.c_uint32),
                ("decay_base", C.c_float), ("step", C.c_uint32), ("lr_factor", C.c_float)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_pcg32_next_float.restype = C.c_float
        _LIB.orc_pcg32_next_uint.restype = C.c_uint32
        _LIB.orc_grid_offsets.restype = C.c_uint32
        _LIB.orc_generate_training_samples.restype = C.c_uint32
        _LIB.orc_compute_loss.restype = C.c_uint32
        _LIB.orc_compute_loss_exposure.restype = C.c_uint32
        _LIB.orc_density_grid_mean.restype = C.c_float
        _LIB.orc_trainer_create.restype = C.c_void_p
        _LIB.orc_trainer_n_params.restype = C.c_uint32
        _LIB.orc_trainer_step.restype = C.c_uint32
        _LIB.orc_trainer_bitfield.restype = C.POINTER(C.c_uint8)
        _LIB.orc_trainer_density_grid.restype = C.POINTER(C.c_float)
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def set_num_threads(n):
    """OpenMP threads of the restatement (0: query). Returns the count in effect."""
    return int(lib().orc_set_num_threads(int(n)))


def pcg32(seed, seq=1):
    r = Pcg32()
    lib().orc_pcg32_seed(C.byref(r), C.c_uint64(seed), C.c_uint64(seq))
    return r


def pcg32_advance(r, delta=1 << 32):
    lib().orc_pcg32_advance(C.byref(r), C.c_int64(delta))


def pcg32_next_uint(r):
    return lib().orc_pcg32_next_uint(C.byref(r))


def pcg32_next_float(r):
    return lib().orc_pcg32_next_float(C.byref(r))


def model(n_levels=16, log2_hashmap_size=19, base_resolution=16, per_level_scale=None, aabb_scale=1):
    if per_level_scale is None:
        # Testbed::reset_network, src/testbed.cu:2313-2325 (float math)
        per_level_scale = float(np.exp(np.log(np.float32(2048.0) * np.float32(aabb_scale) / np.float32(base_resolution)) / np.float32(n_levels - 1), dtype=np.float32))
    m = Model()
    lib().orc_model_init(C.byref(m), n_levels, log2_hashmap_size, base_resolution, C.c_float(per_level_scale))
    return m


def make_images(pixels_list, xforms, fx, fy, cx=0.5, cy=0.5, lens=None):
    """pixels_list: list/array of uint8 (or float16 / float32) [h,w,4]; xforms: [n,3,4] ngp camera matrices; lens: (mode, 7 params) applied to every image. Keeps references alive."""
    n = len(pixels_list)
    arr = (Image * n)()
    keep = []
    for i in range(n):
        px = np.ascontiguousarray(pixels_list[i])
        if px.dtype not in (np.float16, np.float32):  # (float16 / float32 arrays: EImageDataType Half / Float, linear premultiplied colours)
            px = np.ascontiguousarray(px, dtype=np.uint8)
        keep.append(px)
        arr[i].pixels = px.ctypes.data
        arr[i].image_type = 2 if px.dtype == np.float32 else 1 if px.dtype == np.float16 else 0
        arr[i].h, arr[i].w = px.shape[0], px.shape[1]
        arr[i].fx, arr[i].fy, arr[i].cx, arr[i].cy = fx, fy, cx, cy
        if lens is not None:
            arr[i].lens_mode = int(lens[0])
            for k in range(7):
                arr[i].lens_params[k] = float(lens[1][k])
        xf = np.asarray(xforms[i], dtype=np.float32).reshape(3, 4)
        col_major = xf.T.reshape(-1)  # column-major 3x4
        for k in range(12):
            arr[i].xform[k] = float(col_major[k])
    arr._keep = keep
    return arr


def grid_forward(m, grid_half, positions, pos_stride=None, scales=None):
    positions = _f32(positions)
    n = positions.shape[0]
    stride = positions.shape[1] if pos_stride is None else pos_stride
    out = np.empty((n, 2 * m.n_levels), dtype=np.float16)
    sc = _f32(scales) if scales is not None else None
    lib().orc_grid_forward(n, m.n_levels, m.offsets, m.base_resolution, C.c_float(np.log2(np.float32(m.per_level_scale))), _p(sc),
                           _p(np.ascontiguousarray(grid_half)), _p(positions), stride, _p(out))
    return out


def grid_indices(m, level, positions, scales=None):
    positions = _f32(positions)
    n = positions.shape[0]
    idx = np.empty((n, 8), dtype=np.uint32)
    w = np.empty((n, 8), dtype=np.float32)
    sc = _f32(scales) if scales is not None else None
    lib().orc_grid_indices(n, level, m.offsets, m.base_resolution, C.c_float(np.log2(np.float32(m.per_level_scale))), _p(sc),
                           _p(positions), positions.shape[1], _p(idx), _p(w))
    return idx, w


def grid_backward(m, positions, dL_dy, scales=None):
    positions = _f32(positions)
    n = positions.shape[0]
    dL_dy = np.ascontiguousarray(dL_dy, dtype=np.float16)
    grad = np.zeros(m.n_grid_params, dtype=np.float32)
    sc = _f32(scales) if scales is not None else None
    lib().orc_grid_backward(n, m.n_levels, m.offsets, m.base_resolution, C.c_float(np.log2(np.float32(m.per_level_scale))), _p(sc),
                            _p(positions), positions.shape[1], _p(dL_dy), _p(grad))
    return grad


def sh4(dirs):
    dirs = _f32(dirs)
    out = np.empty((dirs.shape[0], 16), dtype=np.float16)
    lib().orc_sh4(dirs.shape[0], _p(dirs), dirs.shape[1], _p(out), 16)
    return out


def mlp_forward(mlp_half, encoded, coords, save=False):
    n = encoded.shape[0]
    encoded = np.ascontiguousarray(encoded, dtype=np.float16)
    coords = _f32(coords)
    out = np.empty((n, 4), dtype=np.float16)
    if save:
        h1 = np.empty((n, 64), np.float16); rin = np.empty((n, 32), np.float16); g1 = np.empty((n, 64), np.float16); g2 = np.empty((n, 64), np.float16)
    else:
        h1 = rin = g1 = g2 = None
    lib().orc_nerf_mlp_forward(n, _p(np.ascontiguousarray(mlp_half, dtype=np.float16)), _p(encoded), _p(coords), _p(out), _p(h1), _p(rin), _p(g1), _p(g2))
    return (out, h1, rin, g1, g2) if save else out


def mlp_backward(mlp_half, encoded, coords, dL_dout):
    n = encoded.shape[0]
    denc = np.empty((n, 32), dtype=np.float16)
    grad = np.empty(MLP_PARAMS, dtype=np.float32)
    lib().orc_nerf_mlp_backward(n, _p(np.ascontiguousarray(mlp_half, dtype=np.float16)), _p(np.ascontiguousarray(encoded, dtype=np.float16)), _p(_f32(coords)),
                                _p(np.ascontiguousarray(dL_dout, dtype=np.float16)), _p(denc), _p(grad))
    return denc, grad


def nerf_inference(m, params_half, coords):
    coords = _f32(coords)
    out = np.empty((coords.shape[0], 4), dtype=np.float16)
    lib().orc_nerf_inference(C.byref(m), _p(np.ascontiguousarray(params_half, dtype=np.float16)), coords.shape[0], _p(coords), _p(out))
    return out


def nerf_density(m, params_half, positions):
    positions = _f32(positions)
    out = np.empty(positions.shape[0], dtype=np.float16)
    lib().orc_nerf_density(C.byref(m), _p(np.ascontiguousarray(params_half, dtype=np.float16)), positions.shape[0], _p(positions), positions.shape[1], _p(out))
    return out


def nerf_forward_backward(m, params_half, coords, dL_dout):
    coords = _f32(coords)
    grad = np.empty(MLP_PARAMS + m.n_grid_params, dtype=np.float32)
    lib().orc_nerf_forward_backward(C.byref(m), _p(np.ascontiguousarray(params_half, dtype=np.float16)), coords.shape[0], _p(coords),
                                    _p(np.ascontiguousarray(dL_dout, dtype=np.float16)), _p(grad))
    return grad


def effective_xform(xf34):
    src = _f32(np.asarray(xf34).reshape(3, 4).T.reshape(-1))
    dst = np.empty(12, dtype=np.float32)
    lib().orc_effective_xform(_p(src), _p(dst))
    return dst.reshape(4, 3).T.copy()


def generate_training_samples(n_rays, aabb6, max_samples, rng, images, bitfield, snap=True, cone_angle=0.0, ray_offset=0, n_rays_global=None):
    aabb6 = _f32(aabb6)
    if n_rays_global is not None:
        ray_indices = np.zeros(n_rays, np.uint32); rays = np.zeros((n_rays, 6), np.float32); numsteps = np.zeros((n_rays, 2), np.uint32)
        coords = np.zeros((max_samples, 7), np.float32); counters = np.zeros(2, np.uint32)
        lib().orc_generate_training_samples_sharded.restype = C.c_uint32
        kept = lib().orc_generate_training_samples_sharded(n_rays, int(ray_offset), int(n_rays_global), _p(aabb6), max_samples, rng, len(images), images,
                                                           _p(np.ascontiguousarray(bitfield)), int(snap), C.c_float(cone_angle), _p(ray_indices), _p(rays), _p(numsteps),
                                                           _p(coords), _p(counters))
        return dict(n_kept=kept, counters=counters, ray_indices=ray_indices, rays=rays, numsteps=numsteps, coords=coords)
    ray_indices = np.zeros(n_rays, np.uint32); rays = np.zeros((n_rays, 6), np.float32); numsteps = np.zeros((n_rays, 2), np.uint32)
    coords = np.zeros((max_samples, 7), np.float32); counters = np.zeros(2, np.uint32)
    kept = lib().orc_generate_training_samples(n_rays, _p(aabb6), max_samples, 0, rng, len(images), images, _p(np.ascontiguousarray(bitfield)),
                                               int(snap), C.c_float(cone_angle), _p(ray_indices), _p(rays), _p(numsteps), _p(coords), _p(counters))
    return dict(n_kept=kept, counters=counters, ray_indices=ray_indices, rays=rays, numsteps=numsteps, coords=coords)


def compute_loss(n_kept, n_rays, aabb6, rng, batch, images, rgbsigma, ray_indices, rays, numsteps, coords_in, mean_density,
                 loss_scale=128.0, background=(0, 0, 0), color_space=1, random_bg=True, linear_colors=False, loss_type=4,
                 rgb_activation=2, density_activation=3, snap=True, near_distance=0.2, exposure=None, want_exposure_gradient=False):
    aabb6 = _f32(aabb6)
    numsteps = np.array(numsteps, dtype=np.uint32, copy=True)
    coords_out = np.zeros((batch, 7), np.float32); dloss = np.zeros((batch, 4), np.float16); loss = np.zeros(n_rays, np.float32)
    bg = _f32(background)
    exposure = None if exposure is None else _f32(exposure).reshape(len(images), 3)
    exposure_gradient = np.zeros((len(images), 3), np.float32) if want_exposure_gradient else None
    total = lib().orc_compute_loss_exposure(n_kept, n_rays, _p(aabb6), 0, rng, batch, C.c_float(loss_scale), _p(bg), color_space, int(random_bg), int(linear_colors),
                                   len(images), images, _p(np.ascontiguousarray(rgbsigma, dtype=np.float16)), _p(np.ascontiguousarray(ray_indices, dtype=np.uint32)),
                                   _p(_f32(rays)), _p(numsteps), _p(_f32(coords_in)), _p(coords_out), _p(dloss), loss_type, _p(loss),
                                   rgb_activation, density_activation, int(snap), C.c_float(mean_density), C.c_float(near_distance),
                                   None if exposure is None else _p(exposure), None if exposure_gradient is None else _p(exposure_gradient))
    return dict(compacted=total, numsteps=numsteps, coords_out=coords_out, dloss=dloss, loss=loss, exposure_gradient=exposure_gradient)


# ---- K19: error-map importance sampling (global state of the restatement: set, call generate_training_samples / compute_loss, clear) ----
import contextlib


@contextlib.contextmanager
def error_sampling(cdf_x_cond_y=None, cdf_y=None, cdf_img=None, error_map=None):
    """Inside the block generate_training_samples / compute_loss / compute_cam_gradient draw pixels from cdf_x_cond_y [n_img, ry, rx] + cdf_y [n_img, ry]
    and images from cdf_img [n_img] (either may be None), and compute_loss deposits every ray's loss into error_map [n_img, ry, rx] (float32, in place)."""
    keep = []
    rx = ry = 0
    if cdf_x_cond_y is not None:
        cdf_x_cond_y = _f32(cdf_x_cond_y); cdf_y = _f32(cdf_y)
        ry, rx = cdf_x_cond_y.shape[1:]
        keep += [cdf_x_cond_y, cdf_y]
    if cdf_img is not None:
        cdf_img = _f32(cdf_img); keep.append(cdf_img)
    lib().orc_set_error_cdf(None if cdf_x_cond_y is None else _p(cdf_x_cond_y), None if cdf_x_cond_y is None else _p(cdf_y),
                            None if cdf_img is None else _p(cdf_img), int(rx), int(ry))
    if error_map is not None:
        assert error_map.dtype == np.float32 and error_map.flags.c_contiguous and error_map.ndim == 3
        lib().orc_set_error_map(_p(error_map), int(error_map.shape[2]), int(error_map.shape[1]))
    try:
        yield
    finally:
        lib().orc_set_error_cdf(None, None, None, 0, 0)
        lib().orc_set_error_map(None, 0, 0)


def construct_cdfs(error_map):
    """construct_cdf_2d + construct_cdf_1d + the host normalisation: error_map [n_img, ry, rx] -> dict(cdf_x_cond_y, cdf_y, image_sums, pmf_img, cdf_img)."""
    em = _f32(error_map)
    n, ry, rx = em.shape
    x = np.zeros_like(em); y = np.zeros((n, ry), np.float32); sums = np.zeros(n, np.float32)
    lib().orc_construct_cdfs(n, ry, rx, _p(em), _p(x), _p(y), _p(sums))
    pmf = np.zeros(n, np.float32); cdf = np.zeros(n, np.float32)
    lib().orc_normalize_image_cdf(n, _p(sums), _p(pmf), _p(cdf))
    return dict(cdf_x_cond_y=x, cdf_y=y, image_sums=sums, pmf_img=pmf, cdf_img=cdf)


def fill_rollover(batch, n_valid, coords, dloss):
    lib().orc_fill_rollover(batch, n_valid, _p(coords), _p(dloss))


def optimizer():
    o = Optimizer()
    lib().orc_optimizer_init(C.byref(o))
    return o


def optimizer_step(o, n_matrix, loss_scale, grad, w_fp32, w_half, w_ema, m1, m2, steps):
    lib().orc_optimizer_step(C.byref(o), w_fp32.shape[0], n_matrix, C.c_float(loss_scale), _p(_f32(grad)), _p(w_fp32), _p(w_half), _p(w_ema), _p(m1), _p(m2), _p(steps))


def mark_untrained(grid, images, clear_visible=True):
    lib().orc_mark_untrained_density_grid(grid.shape[0], _p(grid), len(images), images, int(clear_visible))


def generate_grid_samples(n, rng, step, aabb6, grid_in, n_cascades, thresh):
    pos = np.zeros((n, 3), np.float32); idx = np.zeros(n, np.uint32)
    lib().orc_generate_grid_samples(n, rng, step, _p(_f32(aabb6)), _p(grid_in), _p(pos), _p(idx), n_cascades, C.c_float(thresh))
    return pos, idx


def splat_and_ema(indices, density_half, decay, grid):
    lib().orc_splat_and_ema(indices.shape[0], _p(indices), _p(np.ascontiguousarray(density_half, dtype=np.float16)), grid.shape[0], C.c_float(decay), _p(grid))


def density_grid_mean(grid):
    return lib().orc_density_grid_mean(_p(grid))


def bitfield(n_cascades_used, grid, mean):
    out = np.zeros(GRID_CELLS * 8 // 8, np.uint8)
    lib().orc_bitfield(n_cascades_used, _p(grid), C.c_float(mean), _p(out))
    return out


class RenderConfig(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("fx", C.c_float), ("fy", C.c_float), ("screen_center", C.c_float * 2), ("camera", C.c_float * 12),
                ("spp", C.c_int32), ("snap_to_pixel_centers", C.c_int32), ("aabb", C.c_float * 6), ("render_aabb", C.c_float * 6),
                ("cone_angle_constant", C.c_float), ("min_transmittance", C.c_float), ("near_distance", C.c_float),
                ("rgb_activation", C.c_int32), ("density_activation", C.c_int32), ("train_in_linear_colors", C.c_int32),
                ("color_space", C.c_int32), ("output_srgb", C.c_int32), ("exposure", C.c_float), ("background_color", C.c_float * 4), ("tonemap_curve", C.c_int32)]


def render_config(width, height, fx, fy, camera34, spp=1, snap=False, aabb=(0, 0, 0, 1, 1, 1), cone_angle=0.0, min_transmittance=0.01, near_distance=0.0,
                  rgb_activation=2, density_activation=3, linear_colors=False, color_space=0, output_srgb=False, exposure=0.0, background=(0, 0, 0, 1), tonemap_curve=0):
    """Defaults are the reference's Testbed defaults (testbed.h:547,:725,:846,:853,:875,:889)."""
    c = RenderConfig()
    c.width, c.height, c.fx, c.fy = width, height, fx, fy
    c.screen_center[0] = c.screen_center[1] = 0.5
    cm = np.asarray(camera34, dtype=np.float32).reshape(3, 4).T.reshape(-1)
    for k in range(12):
        c.camera[k] = float(cm[k])
    c.spp, c.snap_to_pixel_centers = spp, int(snap)
    for k in range(6):
        c.aabb[k] = c.render_aabb[k] = float(aabb[k])
    c.cone_angle_constant, c.min_transmittance, c.near_distance = cone_angle, min_transmittance, near_distance
    c.rgb_activation, c.density_activation, c.train_in_linear_colors = rgb_activation, density_activation, int(linear_colors)
    c.color_space, c.output_srgb, c.exposure, c.tonemap_curve = color_space, int(output_srgb), exposure, int(tonemap_curve)
    for k in range(4):
        c.background_color[k] = float(background[k])
    return c


def render_nerf(m, params_half, bitfield, cfg):
    """Classic render of one frame on the CPU: float32 [H][W][4] and the number of samples marched."""
    out = np.zeros((cfg.height, cfg.width, 4), np.float32)
    ns = C.c_uint64(0)
    lib().orc_render_nerf(C.byref(m), _p(np.ascontiguousarray(params_half, dtype=np.float16)), _p(np.ascontiguousarray(bitfield)), C.byref(cfg), _p(out), C.byref(ns))
    return out, int(ns.value)


def grid_offsets_nd(n_dims, n_levels=16, log2_hashmap_size=19, base_resolution=16, per_level_scale=2.0):
    offsets = (C.c_uint32 * 33)()
    lib().orc_grid_offsets_nd.restype = C.c_uint32
    total = lib().orc_grid_offsets_nd(n_dims, n_levels, log2_hashmap_size, base_resolution, C.c_float(per_level_scale), offsets)
    return np.array(offsets[:n_levels + 1], np.uint32), int(total)


def grid_forward_nd(n_dims, offsets, grid_half, positions, per_level_scale, base_resolution=16, scales=None):
    """Hash-grid forward for a 2-D or 3-D input: float16 [n][2 * n_levels]."""
    positions = _f32(positions)
    n, n_levels = positions.shape[0], len(offsets) - 1
    out = np.zeros((n, 2 * n_levels), np.float16)
    off = (C.c_uint32 * 33)(*[int(v) for v in offsets])
    sc = _f32(scales) if scales is not None else None
    lib().orc_grid_forward_nd(n_dims, n, n_levels, off, base_resolution, C.c_float(np.log2(np.float32(per_level_scale))), _p(sc),
                              _p(np.ascontiguousarray(grid_half, np.float16)), _p(positions), positions.shape[1], _p(out))
    return out


def mlp_forward_backward(weights_half, input_half, n_hidden=2, dL_dout=None, want_input_grad=True):
    """FullyFusedMLP 32 -> 64 x n_hidden -> 16 on the CPU: out [n][16]; with dL_dout also (dL_dinput [n][32], grad fp32 [n_params])."""
    w = np.ascontiguousarray(weights_half, np.float16)
    x = np.ascontiguousarray(input_half, np.float16)
    n = x.shape[0]
    out = np.zeros((n, 16), np.float16)
    if dL_dout is None:
        lib().orc_mlp_forward_backward(n_hidden, n, _p(w), _p(x), _p(out), None, None, None)
        return out
    dy = np.ascontiguousarray(dL_dout, np.float16)
    din = np.zeros((n, 32), np.float16) if want_input_grad else None
    grad = np.zeros(w.shape[0], np.float32)
    lib().orc_mlp_forward_backward(n_hidden, n, _p(w), _p(x), _p(out), _p(dy), _p(din), _p(grad))
    return out, din, grad


def loss(kind, predictions_half, targets, loss_scale=128.0):
    """tcnn L2 (kind 0) / MAPE (kind 1) loss on the padded [n][16] fp16 network output: (values fp32 [n][16], gradients fp16 [n][16])."""
    pred = np.ascontiguousarray(predictions_half, np.float16)
    tgt = _f32(targets)
    n, dims = pred.shape[0], tgt.shape[1]
    values = np.zeros((n, 16), np.float32)
    grads = np.zeros((n, 16), np.float16)
    lib().orc_loss(kind, n, dims, C.c_float(loss_scale), _p(pred), _p(tgt), _p(values), _p(grads))
    return values, grads


# ---- input gradients (camera-extrinsics optimisation, K13 / K14; pinned on tests/golden/ref_camera.npz) ----
def grid_input_gradient(m, grid_half, positions, dL_dy, scales=None):
    positions = _f32(positions)
    n = positions.shape[0]
    out = np.zeros((n, 3), np.float32)
    sc = _f32(scales) if scales is not None else None
    lib().orc_grid_input_gradient(n, m.n_levels, m.offsets, m.base_resolution, C.c_float(np.log2(np.float32(m.per_level_scale))), _p(sc),
                                  _p(np.ascontiguousarray(grid_half, np.float16)), _p(positions), positions.shape[1], _p(np.ascontiguousarray(dL_dy, np.float16)), _p(out))
    return out


def sh4_input_gradient(dirs, dL_dy):
    dirs = _f32(dirs)
    out = np.zeros((dirs.shape[0], 3), np.float32)
    lib().orc_sh4_input_gradient(dirs.shape[0], _p(dirs), dirs.shape[1], _p(np.ascontiguousarray(dL_dy, np.float16)), _p(out))
    return out


def nerf_input_gradient(m, params_half, coords, dL_dout):
    coords = _f32(coords)
    out = np.zeros((coords.shape[0], 7), np.float32)
    lib().orc_nerf_input_gradient(C.byref(m), _p(np.ascontiguousarray(params_half, np.float16)), coords.shape[0], _p(coords), _p(np.ascontiguousarray(dL_dout, np.float16)), _p(out))
    return out


def compute_cam_gradient(n_kept, n_rays_total, n_images, aabb6, ray_indices, rays_unnormalized, numsteps, coords, coords_gradient):
    pos = np.zeros((n_images, 3), np.float32); rot = np.zeros((n_images, 3), np.float32)
    lib().orc_compute_cam_gradient(n_kept, n_rays_total, n_images, _p(_f32(aabb6)), _p(np.ascontiguousarray(ray_indices, np.uint32)), _p(_f32(rays_unnormalized)),
                                   _p(np.ascontiguousarray(numsteps, np.uint32)), _p(_f32(coords)), _p(_f32(coords_gradient)), _p(pos), _p(rot))
    return pos, rot


class Mask(C.Structure):  # orc_mask
    _fields_ = [("shape", C.c_int32), ("mode", C.c_int32), ("transform", C.c_float * 16), ("config", C.c_float * 6), ("feather", C.c_float), ("opacity", C.c_float)]


class NerfInstance(C.Structure):
    _fields_ = [("model", C.POINTER(Model)), ("params", C.c_void_p), ("bitfield", C.c_void_p), ("train_aabb", C.c_float * 6), ("aabb_scale", C.c_uint32),
                ("render_aabb", C.c_float * 6), ("transform", C.c_float * 16), ("opacity", C.c_float),
                ("rgb_activation", C.c_int32), ("density_activation", C.c_int32), ("min_transmittance", C.c_float), ("n_masks", C.c_uint32), ("masks", C.POINTER(Mask))]


class BlenderRequest(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("mip", C.c_int32), ("flip_y", C.c_int32), ("camera", C.c_float * 12),
                ("focal_length", C.c_float), ("near_distance", C.c_float), ("color_space", C.c_int32), ("exposure", C.c_float), ("background_color", C.c_float * 4),
                ("camera_model", C.c_int32), ("aperture_size", C.c_float), ("focus_z", C.c_float), ("spherical_quadrilateral", C.c_float * 3),
                ("quadrilateral_hexahedron", C.c_float * 24), ("tonemap_curve", C.c_int32), ("n_masks", C.c_uint32), ("masks", C.POINTER(Mask))]


def mask_array(masks):
    """masks: dicts with shape (0 box, 1 cylinder, 2 sphere), mode (0 add, 1 subtract), transform (4x4 row-major numpy), dims, feather, opacity."""
    if not masks:
        return None
    arr = (Mask * len(masks))()
    for a, m in zip(arr, masks):
        a.shape, a.mode, a.feather, a.opacity = int(m["shape"]), int(m["mode"]), float(m["feather"]), float(m["opacity"])
        t = np.asarray(m["transform"], dtype=np.float32).reshape(4, 4).T.reshape(-1)
        for k in range(16):
            a.transform[k] = float(t[k])
        for k, v in enumerate(m["dims"]):
            a.config[k] = float(v)
    return arr


def blender_render(width, height, camera34, focal_length, nerfs, mip=0, flip_y=False, near_distance=0.0, color_space=1, exposure=0.0, background=(0, 0, 0, 0),
                   camera_model=0, aperture_size=0.0, focus_z=1.0, spherical_quadrilateral=(0, 0, 0), quadrilateral_hexahedron=None, tonemap_curve=0, masks=()):
    """NerfRenderer::render on the CPU (src/nerf_renderer.cu:565-791). nerfs: list of dicts with model, params_half, bitfield, aabb_scale,
    render_aabb (6), transform (4x4 row-major numpy, local -> world), opacity, masks. Returns float32 [H][W][4] and the number of composited samples."""
    rq = BlenderRequest()
    rq.camera_model, rq.aperture_size, rq.focus_z, rq.tonemap_curve = int(camera_model), aperture_size, focus_z, int(tonemap_curve)
    for k in range(3):
        rq.spherical_quadrilateral[k] = float(spherical_quadrilateral[k])
    if quadrilateral_hexahedron is not None:
        for k, v in enumerate(np.asarray(quadrilateral_hexahedron, np.float32).reshape(-1)):
            rq.quadrilateral_hexahedron[k] = float(v)
    request_masks = mask_array(list(masks))
    if request_masks is not None:
        rq.n_masks, rq.masks = len(request_masks), request_masks
    rq.width, rq.height, rq.mip, rq.flip_y = width, height, mip, int(flip_y)
    cm = np.asarray(camera34, dtype=np.float32).reshape(3, 4).T.reshape(-1)
    for k in range(12):
        rq.camera[k] = float(cm[k])
    rq.focal_length, rq.near_distance, rq.color_space, rq.exposure = focal_length, near_distance, color_space, exposure
    for k in range(4):
        rq.background_color[k] = float(background[k])
    arr = (NerfInstance * max(len(nerfs), 1))()
    keep = []
    for i, n in enumerate(nerfs):
        params = np.ascontiguousarray(n["params_half"], dtype=np.float16)
        bits = np.ascontiguousarray(n["bitfield"], dtype=np.uint8)
        keep += [params, bits, n["model"]]
        arr[i].model = C.pointer(n["model"])
        arr[i].params, arr[i].bitfield = _p(params), _p(bits)
        half = 0.5 * n["aabb_scale"]
        arr[i].aabb_scale = n["aabb_scale"]
        t = np.asarray(n.get("transform", np.eye(4)), dtype=np.float32).reshape(4, 4).T.reshape(-1)
        for k in range(16):
            arr[i].transform[k] = float(t[k])
        ra = n.get("render_aabb", (0.5 - half,) * 3 + (0.5 + half,) * 3)
        for k in range(6):
            arr[i].train_aabb[k] = 0.5 - half if k < 3 else 0.5 + half
            arr[i].render_aabb[k] = float(ra[k])
        arr[i].opacity = n.get("opacity", 1.0)
        arr[i].rgb_activation, arr[i].density_activation, arr[i].min_transmittance = 2, 3, 0.01
        ma = mask_array(n.get("masks", []))
        if ma is not None:
            keep.append(ma)
            arr[i].n_masks, arr[i].masks = len(ma), ma
    out = np.zeros((height, width, 4), np.float32)
    ns = C.c_uint64(0)
    lib().orc_blender_render(C.byref(rq), len(nerfs), arr, _p(out), C.byref(ns))
    return out, int(ns.value)


class Trainer:
    """Whole-iteration CPU restatement of Testbed::train for the NeRF mode (oracle/ngp_trainer.cpp)."""

    def __init__(self, images, aabb_scale=1, seed=1337):
        self._images = images
        self._h = C.c_void_p(lib().orc_trainer_create(len(images), images, aabb_scale, seed))
        self.n_params = lib().orc_trainer_n_params(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_trainer_destroy(self._h)
            self._h = None

    def train(self, batch_size):
        stats = np.zeros(4, np.float32)
        lib().orc_trainer_train(self._h, batch_size, _p(stats))
        return dict(loss=float(stats[0]), rays_per_batch=int(stats[1]), measured_batch_size_before_compaction=int(stats[2]), measured_batch_size=int(stats[3]))

    @property
    def training_step(self):
        return lib().orc_trainer_step(self._h)

    def set_level_scales(self, scales):
        lib().orc_trainer_set_level_scales(self._h, _p(_f32(scales)))

    def set_state(self, training_step, rays_per_batch=0, density_grid=None):
        g = _f32(density_grid) if density_grid is not None else None
        lib().orc_trainer_set_state(self._h, int(training_step), int(rays_per_batch), _p(g))

    def set_optimizer_state(self, m1, m2, param_steps, optimizer_step, lr_factor, measured_batch_size_before_compaction, n_rays_total):
        """Teacher forcing (tests/test_gpu_kernels.py::test_training_iteration_teacher_forced)."""
        lib().orc_trainer_set_optimizer_state(self._h, _p(_f32(m1)), _p(_f32(m2)), _p(np.ascontiguousarray(param_steps, np.uint32)), int(optimizer_step), C.c_float(lr_factor),
                                              int(measured_batch_size_before_compaction), int(n_rays_total))

    def params(self):
        w = np.empty(self.n_params, np.float32); h = np.empty(self.n_params, np.float16); e = np.empty(self.n_params, np.float16)
        lib().orc_trainer_get_params(self._h, _p(w), _p(h), _p(e))
        return w, h, e

    def set_params(self, w_fp32):
        lib().orc_trainer_set_params(self._h, _p(_f32(w_fp32)))

    def bitfield(self):
        return np.ctypeslib.as_array(lib().orc_trainer_bitfield(self._h), shape=(GRID_CELLS * 8 // 8,)).copy()

    def density_grid(self, n_cascades=1):
        return np.ctypeslib.as_array(lib().orc_trainer_density_grid(self._h), shape=(GRID_CELLS * n_cascades,)).copy()
