"""`pyngp` surface of the B200-native NeRF hot path.

Mirrors, for the NeRF train/render path only, the pybind11 module of the reference
(src/python_api.cu:306-888): `Testbed(TestbedMode.Nerf)`, `load_training_data`, `reload_network_from_file`,
`train`, `frame`, `render`, and the training/rendering properties scripts/run.py uses. The work is done
by libngpb200.so (hand-written sm_100a CUDA behind the C ABI in include/ngpb.h); this module is a ctypes
binding plus the host-side file parsing (transforms.json, network config JSON) that the reference does
in C++ with nlohmann::json / stb_image. There is no CPU fallback: importing this module without the
built library, or creating a Testbed without a B200, raises.
"""
import ctypes as C
import enum
import json
import math
import os

import numpy as np

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB_PATH = os.path.join(_PKG_DIR, "libngpb200.so")


class TestbedMode(enum.Enum):  # python_api.cu:311-317
    Nerf = 0
    Sdf = 1
    Image = 2
    Volume = 3


class LossType(enum.IntEnum):  # common.h:103-111
    L2 = 0
    L1 = 1
    Mape = 2
    Smape = 3
    Huber = 4
    LogL1 = 5
    RelativeL2 = 6


class NerfActivation(enum.IntEnum):  # common.h:114-119
    None_ = 0
    ReLU = 1
    Logistic = 2
    Exponential = 3


class ColorSpace(enum.IntEnum):  # common.h:129-133
    Linear = 0
    SRGB = 1


# ---- C structs (include/ngpb.h) ---------------------------------------------------------------
NGPB_MAX_LEVELS = 32


class TonemapCurve(enum.IntEnum):  # common.h:136-141 ETonemapCurve
    Identity = 0
    ACES = 1
    Hable = 2
    Reinhard = 3


class CameraModel(enum.IntEnum):  # camera_models.cuh:27-31 ECameraModel (integer values as declared there)
    Perspective = 0
    QuadrilateralHexahedron = 1
    SphericalQuadrilateral = 2


class MaskMode(enum.IntEnum):  # nerf/mask_3D.cuh:16-19
    Add = 0
    Subtract = 1


class MaskShape(enum.IntEnum):  # nerf/mask_3D.cuh:21-26
    Box = 0
    Cylinder = 1
    Sphere = 2
    All = 3


# ---- value types of the Blender render request (python_api.cu:409-538; nerf/render_request.cuh, nerf/nerf_descriptor.cuh, bounding_box.cuh) ----
class BoundingBox:
    def __init__(self, min=None, max=None):
        self.min = np.full(3, np.inf, np.float32) if min is None else np.asarray(min, np.float32).reshape(3).copy()
        self.max = np.full(3, -np.inf, np.float32) if max is None else np.asarray(max, np.float32).reshape(3).copy()

    def center(self):
        return 0.5 * (self.min + self.max)

    def diag(self):
        return self.max - self.min

    def contains(self, p):
        p = np.asarray(p, np.float32)
        return bool(np.all(p >= self.min) and np.all(p <= self.max))

    def enlarge(self, other):
        if isinstance(other, BoundingBox):
            self.min, self.max = np.minimum(self.min, other.min), np.maximum(self.max, other.max)
        else:
            p = np.asarray(other, np.float32)
            self.min, self.max = np.minimum(self.min, p), np.maximum(self.max, p)

    def inflate(self, amount):
        self.min, self.max = self.min - np.float32(amount), self.max + np.float32(amount)

    def intersection(self, other):
        return BoundingBox(np.maximum(self.min, other.min), np.minimum(self.max, other.max))

    def intersects(self, other):
        return not bool(np.any(self.intersection(other).max < self.intersection(other).min))


class DownsampleInfo:
    """DownsampleInfo::MakeFromMip (common.h): every `skip`-th pixel is traced and replicated into a skip x skip block."""

    def __init__(self, resolution, mip):
        self.max_res = (int(resolution[0]), int(resolution[1]))
        self.mip = int(mip)
        self.skip = 1 << self.mip
        self.scaled_res = tuple((r + self.skip - 1) // self.skip for r in self.max_res)

    @staticmethod
    def MakeFromMip(resolution, mip):
        return DownsampleInfo(resolution, mip)


class Mask3D:
    """Mask3D (nerf/mask_3D.cuh:128-257; python_api.cu:446-450): an SDF shape that adds or subtracts visibility inside a NeRF, with a feathered border.
    `transform`: 4x4 shape -> NeRF-local frame (masks of a NerfDescriptor) or shape -> world (masks of a RenderRequest)."""

    def __init__(self, shape, transform, mode, config, feather, opacity):
        self.shape, self.mode = MaskShape(int(shape)), MaskMode(int(mode))
        self.transform = np.asarray(transform, np.float32).reshape(4, 4).copy()
        self.config = [float(v) for v in config] + [0.0] * (6 - len(config))
        self.feather, self.opacity = float(feather), float(opacity)

    @staticmethod
    def Box(dims, transform, mode, feather, opacity):
        d = np.asarray(dims, np.float32).reshape(3)
        return Mask3D(MaskShape.Box, transform, mode, [d[0], d[1], d[2]], feather, opacity)

    @staticmethod
    def Cylinder(radius, height, transform, mode, feather, opacity):
        return Mask3D(MaskShape.Cylinder, transform, mode, [radius, height], feather, opacity)

    @staticmethod
    def Sphere(radius, transform, mode, feather, opacity):
        return Mask3D(MaskShape.Sphere, transform, mode, [radius], feather, opacity)


class RenderModifiers:  # RenderModifiersDescriptor (nerf/render_modifiers_descriptor.cuh; python_api.cu:472-474)
    def __init__(self, masks=()):
        self.masks = list(masks)
        for m in self.masks:
            if not isinstance(m, Mask3D):
                raise TypeError("RenderModifiers: masks must be Mask3D objects")


class Quadrilateral3D:  # camera_models.cuh:33-58
    def __init__(self, tl, tr, bl, br):
        self.tl, self.tr, self.bl, self.br = (np.asarray(v, np.float32).reshape(3).copy() for v in (tl, tr, bl, br))

    @staticmethod
    def Zero():
        return Quadrilateral3D(*([np.zeros(3, np.float32)] * 4))

    def center(self):
        return (self.tl + self.tr + self.bl + self.br) / np.float32(4.0)


class QuadrilateralHexahedronConfig:  # QuadrilateralHexahedron, camera_models.cuh:60-80
    def __init__(self, front, back):
        self.front, self.back = front, back

    @staticmethod
    def Zero():
        return QuadrilateralHexahedronConfig(Quadrilateral3D.Zero(), Quadrilateral3D.Zero())

    def center(self):
        return (self.front.center() + self.back.center()) / np.float32(2.0)


class SphericalQuadrilateralConfig:  # SphericalQuadrilateral, camera_models.cuh:119-135
    def __init__(self, width, height, curvature):
        self.width, self.height, self.curvature = float(width), float(height), float(curvature)

    @staticmethod
    def Zero():
        return SphericalQuadrilateralConfig(0.0, 0.0, 0.0)


class RenderOutputProperties:
    def __init__(self, resolution, ds, spp, color_space, tonemap_curve, exposure, background_color, flip_y):
        self.resolution, self.ds, self.spp = (int(resolution[0]), int(resolution[1])), ds, int(spp)
        self.color_space, self.tonemap_curve, self.exposure = ColorSpace(color_space), TonemapCurve(tonemap_curve), float(exposure)
        self.background_color, self.flip_y = [float(v) for v in background_color], bool(flip_y)


class RenderCameraProperties:
    def __init__(self, transform, model, focal_length, near_distance, aperture_size, focus_z, spherical_quadrilateral=None, quadrilateral_hexahedron=None):
        self.transform = np.asarray(transform, np.float32).reshape(3, 4).copy()
        self.model, self.focal_length, self.near_distance = CameraModel(model), float(focal_length), float(near_distance)
        self.aperture_size, self.focus_z = float(aperture_size), float(focus_z)
        self.spherical_quadrilateral, self.quadrilateral_hexahedron = spherical_quadrilateral, quadrilateral_hexahedron

    def __eq__(self, o):
        return (isinstance(o, RenderCameraProperties) and np.array_equal(self.transform, o.transform) and self.model == o.model
                and (self.focal_length, self.near_distance, self.aperture_size, self.focus_z) == (o.focal_length, o.near_distance, o.aperture_size, o.focus_z))

    __hash__ = None


class NerfDescriptor:
    def __init__(self, snapshot_path_str, aabb, transform, modifiers, opacity):
        self.snapshot_path, self.aabb, self.modifiers, self.opacity = str(snapshot_path_str), aabb, modifiers, float(opacity)
        self.transform = np.asarray(transform, np.float32).reshape(4, 4).copy()


class RenderRequest:
    def __init__(self, output, camera, modifiers, nerfs, aabb):
        self.output, self.camera, self.modifiers, self.nerfs, self.aabb = output, camera, modifiers, list(nerfs), aabb


class MaskStruct(C.Structure):  # ngpb_mask
    _fields_ = [("shape", C.c_int32), ("mode", C.c_int32), ("transform", C.c_float * 16), ("config", C.c_float * 6), ("feather", C.c_float), ("opacity", C.c_float)]


class NerfInstance(C.Structure):  # ngpb_nerf_instance
    _fields_ = [("field", C.c_void_p), ("aabb", C.c_float * 6), ("transform", C.c_float * 16), ("opacity", C.c_float), ("n_masks", C.c_uint32), ("masks", C.POINTER(MaskStruct))]


class BlenderRequest(C.Structure):  # ngpb_blender_request
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("mip", C.c_int32), ("flip_y", C.c_int32), ("camera", C.c_float * 12),
                ("focal_length", C.c_float), ("near_distance", C.c_float), ("color_space", C.c_int32), ("exposure", C.c_float), ("background_color", C.c_float * 4),
                ("camera_model", C.c_int32), ("aperture_size", C.c_float), ("focus_z", C.c_float), ("spherical_quadrilateral", C.c_float * 3),
                ("quadrilateral_hexahedron", C.c_float * 24), ("tonemap_curve", C.c_int32), ("n_masks", C.c_uint32), ("masks", C.POINTER(MaskStruct))]


def _mask_array(masks):
    """ctypes array of ngpb_mask for a list of Mask3D (None for an empty list)."""
    if not masks:
        return None
    arr = (MaskStruct * len(masks))()
    for a, m in zip(arr, masks):
        a.shape, a.mode, a.feather, a.opacity = int(m.shape), int(m.mode), m.feather, m.opacity
        t = m.transform.T.reshape(-1)
        for k in range(16):
            a.transform[k] = float(t[k])
        for k in range(6):
            a.config[k] = m.config[k]
    return arr


class Field:
    """One NeRF loaded from a snapshot for the Blender renderer (NeuralRadianceField, nerf/neural_radiance_field.cuh:153-298)."""

    def __init__(self, params_half, density_grid, aabb_scale, device=0):
        params = np.ascontiguousarray(params_half, np.float16)
        grid = np.ascontiguousarray(density_grid, np.float32)
        h = C.c_void_p()
        check(lib().ngpb_field_create(C.byref(h), int(device), int(aabb_scale), params.ctypes.data_as(C.c_void_p), int(params.shape[0]),
                                      grid.ctypes.data_as(C.c_void_p) if grid.size else None, int(grid.size)))
        self._h, self.aabb_scale = h, int(aabb_scale)

    @staticmethod
    def from_snapshot(path, device=0):
        import msgpack
        if not os.path.exists(path):
            raise RuntimeError(f"Snapshot path {path} does not exist.")
        with open(path, "rb") as f:
            cfg = msgpack.unpackb(f.read(), raw=False, strict_map_key=False)
        if "snapshot" not in cfg:
            raise RuntimeError(f"File {path} does not contain a snapshot.")
        snap = parse_snapshot(cfg)
        _validate_network_config(snap["network_config"])
        return Field(snap["params_half"], snap["density_grid"], snap["aabb_scale"], device)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ngpb_field_destroy(self._h)
            self._h = None


class Grid(C.Structure):
    _fields_ = [("n_levels", C.c_uint32), ("base_resolution", C.c_uint32), ("log2_per_level_scale", C.c_float),
                ("offsets", C.c_uint32 * (NGPB_MAX_LEVELS + 1)), ("scale", C.c_float * NGPB_MAX_LEVELS),
                ("resolution", C.c_uint32 * NGPB_MAX_LEVELS), ("n_pos_dims", C.c_uint32)]


class LensMode(enum.IntEnum):  # common.h ELensMode (python_api.cu:385-390)
    Perspective = 0
    OpenCV = 1
    FTheta = 2
    LatLong = 3


class Image(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("w", C.c_int32), ("h", C.c_int32), ("fx", C.c_float), ("fy", C.c_float),
                ("cx", C.c_float), ("cy", C.c_float), ("xform", C.c_float * 12), ("raw_xform", C.c_float * 12), ("lens_mode", C.c_int32), ("lens_params", C.c_float * 7),
                ("image_type", C.c_int32)]


class HostImage(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("w", C.c_int32), ("h", C.c_int32), ("fx", C.c_float), ("fy", C.c_float),
                ("cx", C.c_float), ("cy", C.c_float), ("xform", C.c_float * 12), ("lens_mode", C.c_int32), ("lens_params", C.c_float * 7), ("image_type", C.c_int32)]


IMAGE_BYTE, IMAGE_HALF, IMAGE_FLOAT = 0, 1, 2  # NGPB_IMAGE_* (EImageDataType)


def _image_array(img):
    """A training image as the C ABI takes it: (contiguous [h][w][4] array, NGPB_IMAGE_*). float32 / float16 arrays keep their type, anything else is RGBA8."""
    px = np.asarray(img)
    if px.dtype == np.float32:
        itype = IMAGE_FLOAT
    elif px.dtype == np.float16:
        itype = IMAGE_HALF
    else:
        px, itype = px.astype(np.uint8, copy=False), IMAGE_BYTE
    px = np.ascontiguousarray(px)
    if px.ndim != 3 or px.shape[2] != 4:
        raise RuntimeError("image should be (H,W,C) where C=4")
    return px, itype


class Rng(C.Structure):
    _fields_ = [("state", C.c_uint64), ("inc", C.c_uint64)]


class ErrorCdf(C.Structure):
    """ngpb_error_cdf (include/ngpb.h): device pointers of the error-map CDFs; null members = uniform sampling."""
    _fields_ = [("cdf_x_cond_y", C.c_void_p), ("cdf_y", C.c_void_p), ("cdf_img", C.c_void_p), ("res_x", C.c_int32), ("res_y", C.c_int32)]


class LossConfig(C.Structure):
    _fields_ = [("loss_scale", C.c_float), ("background_color", C.c_float * 3), ("color_space", C.c_int32), ("random_bg_color", C.c_int32),
                ("linear_colors", C.c_int32), ("loss_type", C.c_int32), ("rgb_activation", C.c_int32), ("density_activation", C.c_int32),
                ("snap_to_pixel_centers", C.c_int32), ("near_distance", C.c_float)]


class Optimizer(C.Structure):
    _fields_ = [("learning_rate", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("epsilon", C.c_float),
                ("l2_reg", C.c_float), ("ema_decay", C.c_float), ("decay_start", C.c_uint32), ("decay_interval", C.c_uint32),
                ("decay_base", C.c_float), ("step", C.c_uint32), ("lr_factor", C.c_float)]


class ModelConfig(C.Structure):  # ngpb_model_config
    _fields_ = [("n_pos_dims", C.c_uint32), ("n_output_dims", C.c_uint32), ("n_levels", C.c_uint32), ("log2_hashmap_size", C.c_uint32), ("base_resolution", C.c_uint32),
                ("per_level_scale", C.c_float), ("desired_resolution", C.c_float), ("loss", C.c_int32), ("use_ema", C.c_int32), ("optimizer", Optimizer), ("seed", C.c_uint32)]


class RenderConfig(C.Structure):  # ngpb_render_config
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("fx", C.c_float), ("fy", C.c_float), ("screen_center", C.c_float * 2), ("camera", C.c_float * 12),
                ("spp", C.c_int32), ("snap_to_pixel_centers", C.c_int32), ("aabb", C.c_float * 6), ("render_aabb", C.c_float * 6),
                ("cone_angle_constant", C.c_float), ("min_transmittance", C.c_float), ("near_distance", C.c_float),
                ("rgb_activation", C.c_int32), ("density_activation", C.c_int32), ("train_in_linear_colors", C.c_int32),
                ("color_space", C.c_int32), ("output_srgb", C.c_int32), ("exposure", C.c_float), ("background_color", C.c_float * 4), ("tonemap_curve", C.c_int32)]


class TrainingState(C.Structure):  # ngpb_training_state
    _fields_ = [("training_step", C.c_uint32), ("rays_per_batch", C.c_uint32), ("measured_batch_size", C.c_uint32), ("measured_batch_size_before_compaction", C.c_uint32),
                ("loss", C.c_float), ("optimizer_step", C.c_uint32), ("learning_rate", C.c_float), ("learning_rate_factor", C.c_float)]


# ---- snapshot container (reference: Testbed::save_snapshot / load_snapshot, src/testbed.cu:3008-3106; tcnn Trainer::serialize, trainer.h:270-310;
# Ema / ExponentialDecay / Adam ::serialize, ema.h:190, exponential_decay.h:136, adam.h:282). nlohmann::json::to_msgpack of the network config with a
# "snapshot" object; binary blobs are msgpack bin. Pure functions, usable without a GPU. ----
SNAPSHOT_FORMAT_VERSION = 1


def build_snapshot(network_config, params_half, density_grid, aabb_scale, aabb, training_step, loss, rays_per_batch, measured_batch_size,
                   measured_batch_size_before_compaction, optimizer=None, dataset_transform=None):
    """Returns the dict the reference serialises. params_half: fp16 inference (EMA) parameters in the reference's flat order; density_grid: float32."""
    snap = {
        "n_params": int(params_half.shape[0]), "params_type": "__half", "params_binary": np.ascontiguousarray(params_half, np.float16).tobytes(),
        "version": SNAPSHOT_FORMAT_VERSION, "density_grid_size": 128,
        "density_grid_binary": np.ascontiguousarray(density_grid, np.float32).astype(np.float16).tobytes(),
        "nerf": {"aabb_scale": int(aabb_scale), "rgb": {"rays_per_batch": int(rays_per_batch), "measured_batch_size": int(measured_batch_size),
                                                        "measured_batch_size_before_compaction": int(measured_batch_size_before_compaction)}},
        "training_step": int(training_step), "loss": float(loss),
        "aabb": {"min": [float(v) for v in aabb[:3]], "max": [float(v) for v in aabb[3:]]}, "bounding_radius": 1.0,
    }
    if dataset_transform is not None:
        # The reference stores the whole NerfDataset under nerf.dataset (json_binding.h:120-160) and needs all of it when the key exists; this build keeps
        # only what set_nerf_camera_matrix needs after load_snapshot, under a key of its own that the reference ignores.
        snap["nerf"]["b200_dataset_transform"] = {"scale": float(dataset_transform[0]), "offset": [float(v) for v in dataset_transform[1]],
                                                  "from_mitsuba": bool(dataset_transform[2]) if len(dataset_transform) > 2 else False}
    if optimizer is not None:
        snap["optimizer"] = {  # Ema -> ExponentialDecay -> Adam
            "weights_ema_binary": np.ascontiguousarray(params_half, np.float16).tobytes(),
            "nested": {"learning_rate": float(optimizer["learning_rate"]), "learning_rate_factor": float(optimizer["learning_rate_factor"]),
                       "nested": {"current_step": int(optimizer["current_step"]), "base_learning_rate": float(optimizer["learning_rate"]),
                                  "first_moments_binary": np.ascontiguousarray(optimizer["first_moments"], np.float32).tobytes(),
                                  "second_moments_binary": np.ascontiguousarray(optimizer["second_moments"], np.float32).tobytes(),
                                  "param_steps_binary": np.ascontiguousarray(optimizer["param_steps"], np.uint32).tobytes()}}}
    cfg = json.loads(json.dumps(network_config))
    cfg["snapshot"] = snap
    return cfg


def _blob(v, dtype):
    if isinstance(v, dict) and "bytes" in v:  # nlohmann's JSON rendering of a binary value
        v = bytes(v["bytes"])
    return np.frombuffer(bytes(v), dtype=dtype).copy()


def parse_snapshot(cfg):
    """Inverse of build_snapshot for files written by the reference or by this module. Raises like the reference on malformed files."""
    if "snapshot" not in cfg:
        raise RuntimeError("File does not contain a snapshot.")
    snap = cfg["snapshot"]
    if snap.get("version", 0) < SNAPSHOT_FORMAT_VERSION:
        raise RuntimeError("Snapshot uses an old format.")
    if snap.get("density_grid_size", 128) != 128:
        raise RuntimeError("Incompatible grid size.")
    ptype = snap.get("params_type", "__half")
    if ptype == "__half":
        params = _blob(snap["params_binary"], np.float16)
    elif ptype == "float":
        params = _blob(snap["params_binary"], np.float32).astype(np.float16)
    else:
        raise RuntimeError("Trainer: snapshot parameters must be of type float of __half")
    out = dict(params_half=params, density_grid=_blob(snap.get("density_grid_binary", b""), np.float16).astype(np.float32),
               aabb_scale=int(snap.get("nerf", {}).get("aabb_scale", 1)), training_step=int(snap.get("training_step", 0)), loss=float(snap.get("loss", 0.0)),
               rgb=snap.get("nerf", {}).get("rgb", {}), aabb=snap.get("aabb"), optimizer=None, dataset_transform=None,
               network_config={k: v for k, v in cfg.items() if k != "snapshot"})
    nerf = snap.get("nerf", {})
    for key in ("dataset", "b200_dataset_transform"):  # a reference snapshot carries its NerfDataset; ours the two numbers
        if isinstance(nerf.get(key), dict) and "scale" in nerf[key] and "offset" in nerf[key]:
            out["dataset_transform"] = (float(nerf[key]["scale"]), tuple(float(v) for v in nerf[key]["offset"]))
            out["dataset_from_mitsuba"] = bool(nerf[key].get("from_mitsuba", False))
            break
    if "optimizer" in snap:
        o = snap["optimizer"]
        decay = o.get("nested", {})
        adam = decay.get("nested", {})
        if "first_moments_binary" in adam:
            out["optimizer"] = dict(current_step=int(adam.get("current_step", 0)), learning_rate=float(decay.get("learning_rate", adam.get("base_learning_rate", 1e-2))),
                                    learning_rate_factor=float(decay.get("learning_rate_factor", 1.0)),
                                    first_moments=_blob(adam["first_moments_binary"], np.float32), second_moments=_blob(adam["second_moments_binary"], np.float32),
                                    param_steps=_blob(adam["param_steps_binary"], np.uint32) if "param_steps_binary" in adam else None)
    return out


# every symbol include/ngpb.h declares (checked by tests/test_abi.py)
EXPORTED_SYMBOLS = [
    "ngpb_last_error", "ngpb_version", "ngpb_check_device", "ngpb_grid_init", "ngpb_hash_encode_forward", "ngpb_hash_encode_backward",
    "ngpb_nerf_mlp_forward", "ngpb_nerf_mlp_workspace_bytes", "ngpb_nerf_mlp_forward_backward", "ngpb_nerf_density_mlp_forward",
    "ngpb_effective_xform", "ngpb_generate_training_samples", "ngpb_compute_loss", "ngpb_optimizer_init", "ngpb_optimizer_step",
    "ngpb_mark_untrained_density_grid", "ngpb_generate_grid_samples", "ngpb_splat_and_ema", "ngpb_update_bitfield", "ngpb_selftest_umma",
    "ngpb_testbed_create", "ngpb_testbed_destroy", "ngpb_testbed_load_training_data", "ngpb_testbed_reset_network", "ngpb_testbed_train",
    "ngpb_testbed_train_n", "ngpb_testbed_loss", "ngpb_testbed_training_step", "ngpb_testbed_stats", "ngpb_testbed_n_params",
    "ngpb_testbed_get_params", "ngpb_testbed_set_params", "ngpb_testbed_get_density_grid", "ngpb_testbed_set_option", "ngpb_testbed_get_option",
    "ngpb_testbed_render", "ngpb_testbed_stream", "ngpb_testbed_stage_times", "ngpb_grid_device_scales", "ngpb_testbed_configure", "ngpb_testbed_set_params_half", "ngpb_testbed_set_density_grid", "ngpb_testbed_get_training_state",
    "ngpb_testbed_set_training_state", "ngpb_testbed_get_optimizer_state", "ngpb_testbed_set_optimizer_state", "ngpb_generate_training_samples_sharded", "ngpb_compute_loss_sharded", "ngpb_nccl_unique_id", "ngpb_testbed_init_data_parallel", "ngpb_render_workspace_bytes", "ngpb_render_nerf", "ngpb_testbed_last_render_ms", "ngpb_generate_training_samples_scratch_bytes", "ngpb_compute_loss_scratch_bytes",
    "ngpb_field_create", "ngpb_field_destroy", "ngpb_blender_render", "ngpb_compute_loss_compact_features", "ngpb_grid_init_nd", "ngpb_mlp_forward", "ngpb_mlp_forward_backward", "ngpb_loss",
    "ngpb_nerf_mlp_forward_backward_sh", "ngpb_nerf_input_gradient", "ngpb_compute_cam_gradient", "ngpb_camera_adam_step", "ngpb_apply_camera_offsets",
    "ngpb_testbed_get_camera_extrinsics", "ngpb_testbed_set_camera_extrinsics", "ngpb_testbed_reset_camera_extrinsics", "ngpb_probe_umma",
    "ngpb_exposure_update", "ngpb_compute_loss_exposure", "ngpb_testbed_get_camera_exposures", "ngpb_testbed_set_camera_exposures",
    "ngpb_generate_training_samples_cdf", "ngpb_compute_loss_error_map", "ngpb_construct_error_cdfs", "ngpb_testbed_get_error_map_pmf",
    "ngpb_testbed_create_empty_dataset", "ngpb_testbed_set_training_image", "ngpb_testbed_set_camera_intrinsics",
    "ngpb_model_create", "ngpb_model_destroy", "ngpb_model_reset", "ngpb_model_n_params", "ngpb_model_training_step", "ngpb_model_loss", "ngpb_model_launches", "ngpb_model_stream",
    "ngpb_model_set_option", "ngpb_model_get_params", "ngpb_model_set_params_half", "ngpb_model_set_training_step", "ngpb_model_train", "ngpb_model_inference",
    "ngpb_model_set_image", "ngpb_model_set_image_rgba8", "ngpb_model_train_image", "ngpb_model_image_mse", "ngpb_model_render_image", "ngpb_model_set_sdf_data",
    "ngpb_model_train_sdf", "ngpb_model_get_training_batch", "ngpb_next_rays_per_batch", "ngpb_ray_shard",
]

_lib = None


def lib():
    """Loads libngpb200.so. Fails loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} is missing: build it with blender-ngp_b200/build.sh (or __graft_entry__.build())")
        l = C.CDLL(_LIB_PATH)
        l.ngpb_last_error.restype = C.c_char_p
        l.ngpb_grid_init.restype = C.c_uint32
        l.ngpb_grid_init_nd.restype = C.c_uint32
        l.ngpb_nerf_mlp_workspace_bytes.restype = C.c_uint64
        l.ngpb_generate_training_samples_scratch_bytes.restype = C.c_uint64
        l.ngpb_compute_loss_scratch_bytes.restype = C.c_uint64
        l.ngpb_render_workspace_bytes.restype = C.c_uint64
        l.ngpb_testbed_last_render_ms.restype = C.c_double
        l.ngpb_testbed_last_render_ms.argtypes = [C.c_void_p]
        l.ngpb_testbed_loss.restype = C.c_float
        l.ngpb_testbed_training_step.restype = C.c_uint32
        l.ngpb_testbed_n_params.restype = C.c_uint32
        l.ngpb_testbed_get_option.restype = C.c_double
        l.ngpb_testbed_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        l.ngpb_testbed_get_option.argtypes = [C.c_void_p, C.c_char_p]
        l.ngpb_effective_xform.restype = None
        l.ngpb_optimizer_init.restype = None
        l.ngpb_testbed_destroy.restype = None
        l.ngpb_field_destroy.restype = None
        l.ngpb_field_destroy.argtypes = [C.c_void_p]
        l.ngpb_testbed_stream.restype = C.c_void_p
        l.ngpb_testbed_stream.argtypes = [C.c_void_p]
        l.ngpb_model_destroy.restype = None
        l.ngpb_model_destroy.argtypes = [C.c_void_p]
        l.ngpb_model_n_params.restype = C.c_uint32
        l.ngpb_model_training_step.restype = C.c_uint32
        l.ngpb_model_loss.restype = C.c_float
        l.ngpb_model_launches.restype = C.c_uint64
        l.ngpb_model_stream.restype = C.c_void_p
        for fn in (l.ngpb_model_n_params, l.ngpb_model_training_step, l.ngpb_model_loss, l.ngpb_model_launches, l.ngpb_model_stream):
            fn.argtypes = [C.c_void_p]
        l.ngpb_model_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        l.ngpb_model_train.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int]
        l.ngpb_model_inference.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
        l.ngpb_model_get_training_batch.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        l.ngpb_model_set_sdf_data.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        l.ngpb_model_set_image.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        l.ngpb_model_set_image_rgba8.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        l.ngpb_model_train_image.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
        l.ngpb_model_train_sdf.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
        l.ngpb_model_set_params_half.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        l.ngpb_model_get_params.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.ngpb_next_rays_per_batch.restype = C.c_uint32
        l.ngpb_next_rays_per_batch.argtypes = [C.c_uint32] * 4
        l.ngpb_ray_shard.restype = None
        l.ngpb_ray_shard.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        l.ngpb_camera_adam_step.restype = None
        l.ngpb_camera_adam_step.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_int]
        l.ngpb_apply_camera_offsets.restype = None
        l.ngpb_apply_camera_offsets.argtypes = [C.c_void_p] * 4
        l.ngpb_testbed_get_camera_extrinsics.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        l.ngpb_testbed_set_camera_extrinsics.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        l.ngpb_testbed_reset_camera_extrinsics.argtypes = [C.c_void_p]
        l.ngpb_exposure_update.restype = None
        l.ngpb_exposure_update.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float]
        l.ngpb_testbed_get_camera_exposures.argtypes = [C.c_void_p, C.c_void_p]
        l.ngpb_testbed_get_error_map_pmf.argtypes = [C.c_void_p, C.c_void_p]
        l.ngpb_testbed_create_empty_dataset.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        l.ngpb_testbed_set_training_image.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        l.ngpb_testbed_set_camera_intrinsics.argtypes = [C.c_void_p, C.c_uint32] + [C.c_float] * 8
        l.ngpb_construct_error_cdfs.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32] + [C.c_void_p] * 5
        l.ngpb_testbed_set_camera_exposures.argtypes = [C.c_void_p, C.c_void_p]
        _lib = l
    return _lib


def free_temporary_memory():
    """pyngp.free_temporary_memory (python_api.cu:309): tcnn's stream-ordered arena has no counterpart here -- workspaces are owned by their Testbed
    and released with it -- so this only synchronises nothing and returns."""
    return None


def check(status):
    if status != 0:
        raise RuntimeError(lib().ngpb_last_error().decode() or f"ngpb error {status}")


def grid_init(n_levels=16, log2_hashmap_size=19, base_resolution=16, per_level_scale=None, aabb_scale=1, device_scales=False, n_pos_dims=3, desired_resolution=2048.0):
    """ngpb_grid for the given hash-grid config; device_scales=True replaces the level scales by the device-evaluated ones (needs a GPU).
    desired_resolution: finest level over the unit cube (Testbed::reset_network, src/testbed.cu:2313-2325): 2048 for NeRF, max(image resolution) / 2 for
    the neural-image model."""
    if per_level_scale is None:
        per_level_scale = float(np.exp(np.log(np.float32(desired_resolution) * np.float32(aabb_scale) / np.float32(base_resolution)) / np.float32(n_levels - 1), dtype=np.float32))
    g = Grid()
    entries = lib().ngpb_grid_init_nd(C.byref(g), n_pos_dims, n_levels, log2_hashmap_size, base_resolution, C.c_float(per_level_scale))
    if device_scales:
        check(lib().ngpb_grid_device_scales(None, C.byref(g)))
    return g, entries


# ---- network config (reference: src/testbed.cu:77-161 load_network_config with `parent` inheritance) ----
def _strip_json_comments(text):
    out, i, in_str = [], 0, False
    while i < len(text):
        ch = text[i]
        if in_str:
            out.append(ch)
            if ch == "\\":
                out.append(text[i + 1]); i += 1
            elif ch == '"':
                in_str = False
        elif ch == '"':
            in_str = True; out.append(ch)
        elif text.startswith("//", i):
            while i < len(text) and text[i] != "\n":
                i += 1
            continue
        elif text.startswith("/*", i):
            i = text.find("*/", i) + 2
            continue
        else:
            out.append(ch)
        i += 1
    return "".join(out)


def _merge_patch(base, patch):  # nlohmann merge_patch (RFC 7386), testbed.cu:77-88
    if not isinstance(patch, dict):
        return patch
    if not isinstance(base, dict):
        base = {}
    out = dict(base)
    for k, v in patch.items():
        if v is None:
            out.pop(k, None)
        else:
            out[k] = _merge_patch(out.get(k), v)
    return out


def load_network_config(path):
    with open(path) as f:
        cfg = json.loads(_strip_json_comments(f.read()))
    while "parent" in cfg:
        parent_path = os.path.join(os.path.dirname(path), cfg.pop("parent"))
        with open(parent_path) as f:
            parent = json.loads(_strip_json_comments(f.read()))
        path = parent_path
        cfg = _merge_patch(parent, cfg)
        if "parent" in parent and "parent" not in cfg:
            cfg["parent"] = parent["parent"]
    return cfg


BASE_NETWORK_CONFIG = {  # configs/nerf/base.json of the reference
    "loss": {"otype": "Huber"},
    "optimizer": {"otype": "Ema", "decay": 0.95, "nested": {"otype": "ExponentialDecay", "decay_start": 20000, "decay_interval": 10000, "decay_base": 0.33,
                  "nested": {"otype": "Adam", "learning_rate": 1e-2, "beta1": 0.9, "beta2": 0.99, "epsilon": 1e-15, "l2_reg": 1e-6}}},
    "encoding": {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 19, "base_resolution": 16},
    "network": {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 1},
    "dir_encoding": {"otype": "Composite", "nested": [{"n_dims_to_encode": 3, "otype": "SphericalHarmonics", "degree": 4}, {"otype": "Identity"}]},
    "rgb_network": {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 2},
}


def _validate_network_config(cfg):
    """The kernels are specialised for the architecture of configs/nerf/base.json; anything else is refused loudly."""
    def req(section, key, value):
        got = cfg.get(section, {}).get(key, BASE_NETWORK_CONFIG[section].get(key))
        same = (str(got).lower() == str(value).lower()) if isinstance(value, str) else (got == value)
        if not same:
            raise RuntimeError(f"unsupported network config: {section}.{key} = {got!r} (this build implements {value!r})")
    req("encoding", "otype", "HashGrid"); req("encoding", "n_levels", 16); req("encoding", "n_features_per_level", 2)
    req("encoding", "base_resolution", 16)
    t = cfg.get("encoding", {}).get("log2_hashmap_size", 19)
    if not (isinstance(t, int) and 14 <= t <= 24):
        raise RuntimeError(f"unsupported network config: encoding.log2_hashmap_size = {t!r} (this build implements 14..24)")
    for s, hidden in (("network", 1), ("rgb_network", 2)):
        req(s, "otype", "FullyFusedMLP"); req(s, "activation", "ReLU"); req(s, "output_activation", "None")
        req(s, "n_neurons", 64); req(s, "n_hidden_layers", hidden)
    nested = cfg.get("dir_encoding", {}).get("nested", [{}])
    if str(cfg.get("dir_encoding", {}).get("otype", "Composite")).lower() != "composite" or str(nested[0].get("otype", "")).lower() != "sphericalharmonics" or nested[0].get("degree", 4) != 4:
        raise RuntimeError("unsupported network config: dir_encoding must be Composite[SphericalHarmonics degree 4, Identity]")
    opt = cfg.get("optimizer", {})
    if str(opt.get("otype", "")).lower() != "ema" or str(opt.get("nested", {}).get("otype", "")).lower() != "exponentialdecay" or \
            str(opt.get("nested", {}).get("nested", {}).get("otype", "")).lower() != "adam":
        raise RuntimeError("unsupported network config: optimizer must be Ema(ExponentialDecay(Adam))")


# ---- transforms.json loader (reference: src/nerf_loader.cu:197-747), host side ------------------------
def nerf_matrix_to_ngp(c2w, scale, offset, from_mitsuba=False):
    """NerfDataset::nerf_matrix_to_ngp (nerf_loader.h:113-132, scale_columns = false). Mitsuba-convention datasets flip columns 0 and 2 instead of cycling
    the axes."""
    m = np.array(c2w, dtype=np.float32)[:3, :4].copy()
    m[:, 1] *= -1
    m[:, 2] *= -1
    m[:, 3] = m[:, 3] * np.float32(scale) + np.asarray(offset, dtype=np.float32)
    if from_mitsuba:
        m[:, 0] *= -1
        m[:, 2] *= -1
        return m
    return m[[1, 2, 0], :].copy()


def ngp_matrix_to_nerf(m, scale, offset, from_mitsuba=False):
    """NerfDataset::ngp_matrix_to_nerf (nerf_loader.h:134-151, scale_columns = false): the inverse of nerf_matrix_to_ngp."""
    m = np.array(m, dtype=np.float32)[:3, :4]
    if from_mitsuba:
        r = m.copy()
        r[:, 0] *= -1
        r[:, 2] *= -1
    else:
        r = m[[2, 0, 1], :].copy()  # cycle axes xyz -> yzx back
    r[:, 1] *= -1
    r[:, 2] *= -1
    r[:, 3] = (r[:, 3] - np.asarray(offset, dtype=np.float32)) / np.float32(scale)
    return r


def _apply_byte_image_rules(img, base, file_path, path, white_transparent, black_transparent):
    """What ngp::load_nerf does to an 8-bit frame besides decoding it (nerf_loader.cu:584-620 and convert_rgba32, :58-81): a separate alpha image
    `<file_path>.alpha.<ext>` (its red channel, sRGB -> linear, replaces the alpha channel), a `dynamic_mask_<name>.png` next to the frame (pixels where its
    red channel is non-zero become 0x00FF00FF, which the training kernels skip), and the dataset's "white_transparent" / "black_transparent" switches."""
    from PIL import Image as PILImage
    ext = os.path.splitext(path)[1][1:]
    alpha_path = os.path.join(base, f"{file_path}.alpha.{ext}")
    if os.path.exists(alpha_path):
        try:
            a = np.asarray(PILImage.open(alpha_path).convert("RGBA"), dtype=np.uint8)
        except OSError:
            raise RuntimeError("Could not load alpha image " + alpha_path)
        if a.shape[:2] != img.shape[:2]:
            raise RuntimeError(f"Alpha image {alpha_path} has wrong resolution.")
        x = a[..., 0].astype(np.float32) * np.float32(1.0 / 255.0)
        lin = np.where(x <= np.float32(0.04045), x / np.float32(12.92), np.power((x + np.float32(0.055)) / np.float32(1.055), np.float32(2.4)))
        img[..., 3] = (np.float32(255.0) * lin.astype(np.float32)).astype(np.uint8)  # (truncating cast, as the reference's (uint8_t))
    mask_path = os.path.join(os.path.dirname(path), f"dynamic_mask_{os.path.splitext(os.path.basename(path))[0]}.png")
    if os.path.exists(mask_path):
        try:
            m = np.asarray(PILImage.open(mask_path).convert("RGBA"), dtype=np.uint8)
        except OSError:
            raise RuntimeError(f"Dynamic mask {mask_path} could not be loaded.")
        if m.shape[:2] != img.shape[:2]:
            raise RuntimeError(f"Dynamic mask {mask_path} has wrong resolution.")
        img[m[..., 0] != 0] = np.array([0xFF, 0x00, 0xFF, 0x00], np.uint8)  # 0x00FF00FF, "hot pink"
    if white_transparent:  # (the NSVF datasets' "white = transparent")
        img[np.all(img[..., :3] == 255, axis=-1), 3] = 0
    if black_transparent:
        img[np.all(img[..., :3] == 0, axis=-1), 3] = 0
    return img


def load_exr_float(path):
    """load_exr (src/tinyexr_wrapper.cu:121-135, tinyexr's LoadEXR): an EXR file as [h][w][4] float32 R, G, B, A (A = 1 when the file has none). Decoded with
    OpenCV (the reference uses tinyexr); a single-channel file is taken as grey."""
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")  # (OpenCV ships its EXR codec switched off)
    try:
        import cv2
    except ImportError as e:
        raise RuntimeError("EXR images need OpenCV's EXR codec: " + str(e))
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise RuntimeError("Failed to load EXR image: " + path)
    img = np.asarray(img, dtype=np.float32)
    if img.ndim == 2:
        img = np.repeat(img[..., None], 3, axis=2)
    out = np.empty(img.shape[:2] + (4,), np.float32)
    out[..., :3] = img[..., 2::-1] if img.shape[2] >= 3 else np.repeat(img[..., :1], 3, axis=2)  # OpenCV hands out B, G, R
    out[..., 3] = img[..., 3] if img.shape[2] >= 4 else 1.0
    return out


def _load_exr_half(path, fix_premult):
    """load_exr_to_gpu + interleave_and_cast_kernel (src/tinyexr_wrapper.cu:41-55,:137-227): an EXR frame as [h][w][4] halfs -- R, G, B as stored (times alpha
    when the dataset says "fix_premult"), A or 1 -- i.e. EImageDataType::Half."""
    img = load_exr_float(path)
    if fix_premult:
        img[..., :3] *= img[..., 3:4]
    return img.astype(np.float16)


def _fov_to_focal_length(resolution, degrees):  # common_device.cuh:473
    return 0.5 * resolution / math.tan(0.5 * degrees * math.pi / 180.0)


def _read_focal_length(js, res, focal):
    """read_focal_length (nerf_loader.cu:271-299): `<axis>_fov` (degrees) before `fl_<axis>` before `camera_angle_<axis>` (radians); x alone sets both
    axes. Returns the updated (fx, fy) or None when `js` carries no focal information."""
    def one(resolution, axis):
        if axis + "_fov" in js:
            return _fov_to_focal_length(resolution, float(js[axis + "_fov"]))
        if "fl_" + axis in js:
            return float(js["fl_" + axis])
        if "camera_angle_" + axis in js:
            return _fov_to_focal_length(resolution, float(js["camera_angle_" + axis]) * 180.0 / math.pi)
        return 0.0
    x_fl, y_fl = one(res[0], "x"), one(res[1], "y")
    if x_fl != 0:
        return (x_fl, y_fl if y_fl != 0 else x_fl)
    if y_fl != 0:
        return (y_fl, y_fl)
    return None


def _read_lens(js, lens, pp):
    """read_lens (nerf_loader.cu:197-269): OpenCV k1, k2, p1, p2 (any non-zero selects the model), f-theta polynomial, lat-long; `cx` / `cy` set the principal
    point. `lens` = [mode, params(7)] and `pp` are updated in place; an inner (per-frame) block without lens keys keeps the outer (dataset) lens."""
    mode = LensMode.Perspective
    for k, key in enumerate(("k1", "k2", "p1", "p2")):
        if key in js:
            lens[1][k] = float(js[key])
            if lens[1][k] != 0.0:
                mode = LensMode.OpenCV
    if "cx" in js:
        pp[0] = float(js["cx"]) / float(js["w"])
    if "cy" in js:
        pp[1] = float(js["cy"]) / float(js["h"])
    if "rolling_shutter" in js and any(float(v) != 0.0 for v in js["rolling_shutter"]):
        raise RuntimeError("rolling-shutter cameras are outside the built scope")
    if "ftheta_p0" in js:
        for k in range(5):
            lens[1][k] = float(js[f"ftheta_p{k}"])
        lens[1][5], lens[1][6] = float(js["w"]), float(js["h"])
        mode = LensMode.FTheta
    if "latlong" in js:
        mode = LensMode.LatLong
    if mode != LensMode.Perspective:
        lens[0] = mode


def byte_image_to_half(img):
    """from_rgba32<__half> (common_device.cuh:563-590): RGBA8 sRGB with straight alpha -> halfs, linear colours times alpha; a masked pixel (0x00FF00FF) becomes
    -1 in all four channels. (What the loader keeps when the training images are sharpened.)"""
    img = np.ascontiguousarray(img, np.uint8)
    x = img.astype(np.float32) * np.float32(1.0 / 255.0)
    lin = np.where(x[..., :3] <= np.float32(0.04045), x[..., :3] / np.float32(12.92), np.power((x[..., :3] + np.float32(0.055)) / np.float32(1.055), np.float32(2.4))).astype(np.float32)
    out = np.empty(img.shape, np.float16)
    out[..., :3] = (lin * x[..., 3:4]).astype(np.float16)
    out[..., 3] = x[..., 3].astype(np.float16)
    out[img.view(np.uint32)[..., 0] == 0x00FF00FF] = np.float16(-1.0)
    return out


def sharpen_image(img, amount):
    """The `sharpen` kernel of NerfDataset::set_training_image (nerf_loader.cu:102-125, :805-826) on a Half / Float image [h][w][4]: centre weight
    4 + 1 / amount minus the four neighbours (by flat index i - 1, i - w, i + 1, i + w; below zero clamps to pixel 0, past the end wraps around), divided by
    1 / amount, clamped at zero -- all four channels, alpha included."""
    px = np.ascontiguousarray(img)
    h, w = px.shape[:2]
    n = h * w
    flat = px.reshape(n, 4).astype(np.float32)
    center_w = np.float32(4.0) + np.float32(1.0) / np.float32(amount)
    inv_totalw = np.float32(1.0) / (center_w - np.float32(4.0))
    i = np.arange(n, dtype=np.int64)
    acc = flat * center_w
    for j in (np.maximum(i - 1, 0), np.maximum(i - w, 0), np.where(i + 1 >= n, i + 1 - n, i + 1), np.where(i + w >= n, i + w - n, i + w)):
        acc = acc - flat[j]
    return np.maximum(np.float32(0.0), acc * inv_totalw).astype(px.dtype).reshape(h, w, 4)


def load_transforms(path, sharpen_amount=0.0):
    """ngp::load_nerf for pinhole RGBA datasets (src/nerf_loader.cu:300-747, Testbed::load_nerf src/testbed_nerf.cu:2735-2758): a transforms json or a
    directory, of which EVERY *.json is loaded (as the reference does: pass the file to train on one split). Frames sorted by file_path, `n_frames` and
    sharpness culling, focal-length key precedence, per-frame focal / principal-point overrides, `scale` / `offset` / `aabb_scale` of the last json.
    Returns decoded images (RGBA8; halfs for EXR frames and for every frame when `sharpen_amount` -- nerf.sharpen, or the dataset's "sharpen" key -- is
    positive), ngp-convention camera matrices and per-image intrinsics (`fx`, `fy`, `cx`, `cy` are scalars when all images agree)."""
    from PIL import Image as PILImage
    if os.path.isdir(path):
        json_paths = sorted(os.path.join(path, p) for p in os.listdir(path) if p.lower().endswith(".json") and os.path.isfile(os.path.join(path, p)))
    elif path.lower().endswith(".json"):
        json_paths = [path]
    else:
        raise RuntimeError("NeRF data path must either be a json file or a directory containing json files.")
    if not json_paths:
        raise RuntimeError("Cannot load NeRF data from an empty set of paths.")
    scale, offset, aabb_scale = 1.0, [0.0, 0.0, 0.0], 1  # NERF_SCALE = 1.0 and zero offset in this fork (nerf_loader.h:28, nerf_loader.cu:406-407)
    images, xforms, fxs, fys, cxs, cys, lenses = [], [], [], [], [], [], []
    per_json = []
    for jp in json_paths:
        with open(jp) as f:
            meta = json.loads(_strip_json_comments(f.read()))
        if not isinstance(meta.get("frames"), list):
            continue  # "does not contain any frames. Skipping."
        base = os.path.dirname(jp)
        frames = sorted(meta["frames"], key=lambda fr: fr["file_path"])  # :356-358
        if "n_frames" in meta:
            frames = frames[:min(len(frames), int(meta["n_frames"]))]
        if frames and "sharpness" in frames[0]:  # kill frames blurrier than their neighbours (:365-390)
            threshold, kept = float(meta.get("sharpness_discard_threshold", 0.0)), []
            for i, fr in enumerate(frames):
                a, b = max(0, i - 3), min(i + 3, len(frames) - 1)
                mean = sum(float(frames[k]["sharpness"]) for k in range(a, b)) / (b - a) if b > a else 0.0
                fr = dict(fr, file_path=fr["file_path"].replace("\\", "/"))
                if os.path.exists(os.path.join(base, fr["file_path"])) and float(fr["sharpness"]) > threshold * mean:
                    kept.append(fr)
            frames = kept
        per_json.append((meta, base, frames))
    if sum(len(fr) for _, _, fr in per_json) == 0:
        raise RuntimeError("No training images were found for NeRF training!")
    # a file with "normal_mts_args" makes the whole dataset Mitsuba-convention (nerf_loader.cu:442-453): other default scale / offset, which the file's own
    # "scale" / "offset" still override, and no axis cycling in the camera matrices
    from_mitsuba = any("normal_mts_args" in meta for meta, _, _ in per_json)
    fix_premult, is_hdr = False, False
    for meta, base, frames in per_json:
        if "fix_premult" in meta:  # (:446-448)
            fix_premult = bool(meta["fix_premult"])
        if "sharpen" in meta:  # (:460-462)
            sharpen_amount = float(meta["sharpen"])
        if "normal_mts_args" in meta:
            scale = float(np.float32(0.66))
            offset = [float(np.float32(0.25) * np.float32(0.66))] * 3
        scale = float(meta.get("scale", scale))
        aabb_scale = int(meta.get("aabb_scale", aabb_scale))
        if "offset" in meta:  # an array, or one number for all three axes (:499-504)
            offset = [float(v) for v in meta["offset"]] if isinstance(meta["offset"], (list, tuple)) else [float(meta["offset"])] * 3
        if "aabb" in meta:  # [[min], [max]]: isotropic scale + translation that fit the box into the unit cube, centred (:506-511)
            lo, hi = np.asarray(meta["aabb"][0], np.float32), np.asarray(meta["aabb"][1], np.float32)
            length = max(np.float32(0.000001), np.abs(hi - lo).max())
            scale_f = np.float32(1.0) / np.float32(length)
            scale = float(scale_f)
            offset = [float(v) for v in ((hi + lo) * np.float32(0.5)) * -scale_f + np.float32(0.5)]
        white_transparent, black_transparent = bool(meta.get("white_transparent", False)), bool(meta.get("black_transparent", False))  # (:464-470)
        lens_json, pp_json = [LensMode.Perspective, [0.0] * 7], [0.5, 0.5]
        _read_lens(meta, lens_json, pp_json)
        for fr in frames:
            p = os.path.join(base, fr["file_path"])
            if os.path.splitext(p)[1] == "":
                if os.path.exists(p + ".png"):
                    p = p + ".png"
                elif os.path.exists(p + ".exr"):
                    p = p + ".exr"
                else:
                    raise RuntimeError("Could not find image file: " + p + ".exr")
            if p.lower().endswith(".exr"):  # HDR frames are kept as halfs, linear (nerf_loader.cu:573-577)
                img = _load_exr_half(p, fix_premult)
                is_hdr = True
            else:
                try:
                    img = np.array(PILImage.open(p).convert("RGBA"), dtype=np.uint8)
                except OSError as e:
                    raise RuntimeError("Could not open image file: " + str(e))
                img = _apply_byte_image_rules(img, base, fr["file_path"], p, white_transparent, black_transparent)
            h, w = img.shape[:2]
            focal = _read_focal_length(meta, (w, h), None)
            focal_frame = _read_focal_length(fr, (w, h), focal)
            focal = focal_frame if focal_frame is not None else focal
            if focal is None:
                raise RuntimeError("Couldn't read fov.")
            lens, pp = [lens_json[0], list(lens_json[1])], list(pp_json)
            _read_lens(fr, lens, pp)  # per-frame lens / principal point override the dataset's (:640-643)
            lenses.append((int(lens[0]), lens[1]))
            images.append(np.ascontiguousarray(img))
            # transform_matrix_start / _end (nerf_loader.cu:517-518, :690-700): a camera that moves while the frame is exposed needs K1's per-pixel transform
            # interpolation, which is not built; a start matrix alone (end defaults to it) is an ordinary frame
            m_start = fr["transform_matrix_start"] if "transform_matrix_start" in fr else fr.get("transform_matrix")
            if m_start is None:
                raise RuntimeError(f"frame {fr.get('file_path')!r} has no transform_matrix")
            if "transform_matrix_end" in fr and not np.array_equal(np.asarray(fr["transform_matrix_end"], np.float64), np.asarray(m_start, np.float64)):
                raise RuntimeError("frames with different transform_matrix_start / transform_matrix_end (rolling shutter, motion blur) are outside the built scope")
            xforms.append(nerf_matrix_to_ngp(m_start, scale, offset, from_mitsuba))
            fxs.append(float(focal[0])); fys.append(float(focal[1])); cxs.append(pp[0]); cys.append(pp[1])
    if sharpen_amount > 0.0:  # applied to every image once all files are read (:732, :805-826); 8-bit frames become halfs first
        images = [sharpen_image(byte_image_to_half(im) if im.dtype == np.uint8 else im, sharpen_amount) for im in images]
    uniform = lambda v: v[0] if all(x == v[0] for x in v) else list(v)
    return dict(images=images, xforms=np.stack(xforms), fx=uniform(fxs), fy=uniform(fys), cx=uniform(cxs), cy=uniform(cys), aabb_scale=aabb_scale,
                scale=scale, offset=offset, lenses=lenses, from_mitsuba=from_mitsuba, is_hdr=is_hdr)


class _CountInt(int):
    """A count the reference exposes as a method (`testbed.n_params()`, python_api.cu:612-613) and earlier versions of this module as a property: both
    spellings work."""

    def __call__(self):
        return int(self)


def _unbuilt(owner, name, default):
    """A reference option this path does not build: it reads as the reference's default and refuses any other value instead of being silently ignored."""
    def setter(s, v):
        same = (list(v) == list(default)) if isinstance(default, (list, tuple)) else (v == default)
        if not same:
            raise RuntimeError(f"{owner}.{name} = {v!r}: not built in this library (only the default {default!r})")
    return property(lambda s: default, setter)


class _Training:
    """`testbed.nerf.training.*` (python_api.cu:744-852): the properties this path reads."""
    def __init__(self, tb):
        self._tb = tb

    def _bool_prop(name):
        return property(lambda s: bool(s._tb._get(name)), lambda s, v: s._tb._set(name, 1.0 if v else 0.0))

    def _unbuilt_prop(name, default):
        """Reference options (python_api.cu:806-827) this path does not build: they read as the reference's default and refuse any other value."""
        def setter(s, v):
            if v != default:
                raise RuntimeError(f"nerf.training.{name} = {v!r}: not built in this library (only the default {default!r})")
        return property(lambda s: default, setter)

    optimize_distortion = _unbuilt_prop("optimize_distortion", False)
    optimize_focal_length = _unbuilt_prop("optimize_focal_length", False)
    optimize_extra_dims = _unbuilt_prop("optimize_extra_dims", False)
    include_sharpness_in_error = _unbuilt_prop("include_sharpness_in_error", False)
    depth_supervision_lambda = _unbuilt_prop("depth_supervision_lambda", 0.0)
    random_bg_color = _bool_prop("random_bg_color")
    linear_colors = _bool_prop("linear_colors")
    snap_to_pixel_centers = _bool_prop("snap_to_pixel_centers")
    loss_type = property(lambda s: LossType(int(s._tb._get("loss_type"))), lambda s, v: s._tb._set("loss_type", int(v)))
    near_distance = property(lambda s: s._tb._get("near_distance"), lambda s, v: s._tb._set("near_distance", float(v)))
    density_grid_decay = property(lambda s: s._tb._get("density_grid_decay"), lambda s, v: s._tb._set("density_grid_decay", float(v)))

    # camera-extrinsics optimisation (python_api.cu:811-844; K13 / K14)
    optimize_extrinsics = _bool_prop("optimize_extrinsics")
    extrinsic_l2_reg = property(lambda s: s._tb._get("extrinsic_l2_reg"), lambda s, v: s._tb._set("extrinsic_l2_reg", float(v)))
    extrinsic_learning_rate = property(lambda s: s._tb._get("extrinsic_learning_rate"), lambda s, v: s._tb._set("extrinsic_learning_rate", float(v)))
    n_steps_between_cam_updates = property(lambda s: int(s._tb._get("n_steps_between_cam_updates")), lambda s, v: s._tb._set("n_steps_between_cam_updates", int(v)))
    n_steps_since_cam_update = property(lambda s: int(s._tb._get("n_steps_since_cam_update")))

    def get_camera_extrinsics(self, frame_idx):
        """Training::get_camera_extrinsics (src/testbed_nerf.cu:2590-2595): the frame's CURRENT training transform (dataset transform + learned offsets),
        as a NeRF-convention 3x4 camera-to-world matrix. Out-of-range indices return the identity, as the reference does."""
        tb = self._tb
        out = np.zeros(12, np.float32)
        if lib().ngpb_testbed_get_camera_extrinsics(tb._h, int(frame_idx) & 0xFFFFFFFF, out.ctypes.data, None, None) != 0:
            return np.eye(4, dtype=np.float32)[:3]
        return ngp_matrix_to_nerf(out.reshape(4, 3).T, tb._dataset_scale, tb._dataset_offset, tb._from_mitsuba)

    def get_camera_offsets(self, frame_idx):
        """(position offset, rotation offset as angle-axis) of a frame: cam_pos_offset[i].variable() / cam_rot_offset[i].variable() (testbed.h:637-640)."""
        pos, rot = np.zeros(3, np.float32), np.zeros(3, np.float32)
        check(lib().ngpb_testbed_get_camera_extrinsics(self._tb._h, int(frame_idx), None, pos.ctypes.data, rot.ctypes.data))
        return pos, rot

    def set_camera_extrinsics(self, frame_idx, camera_to_world, convert_to_ngp=True):
        """Training::set_camera_extrinsics (:2539): replaces the dataset transform of a training frame."""
        tb = self._tb
        m = np.asarray(camera_to_world, dtype=np.float32)[:3, :4]
        if convert_to_ngp:
            m = nerf_matrix_to_ngp(m, tb._dataset_scale, tb._dataset_offset, tb._from_mitsuba)
        flat = np.ascontiguousarray(m.T, np.float32).reshape(-1)  # 3x4 column-major
        check(lib().ngpb_testbed_set_camera_extrinsics(tb._h, int(frame_idx), flat.ctypes.data))

    def reset_camera_extrinsics(self):
        check(lib().ngpb_testbed_reset_camera_extrinsics(self._tb._h))

    # per-image exposure optimisation (python_api.cu:813,:827; train_nerf :3105-3131)
    optimize_exposure = _bool_prop("optimize_exposure")
    exposure_l2_reg = property(lambda s: s._tb._get("exposure_l2_reg"), lambda s, v: s._tb._set("exposure_l2_reg", float(v)))

    def get_camera_exposures(self):
        """The learned per-image exposures in stops, [n_images, 3] (cam_exposure[i].variable(), testbed.h:632). The reference only plots these in
        its GUI (src/testbed.cu:879-885); this accessor is an addition."""
        out = np.zeros((int(self._tb._get("n_images")), 3), np.float32)
        check(lib().ngpb_testbed_get_camera_exposures(self._tb._h, out.ctypes.data))
        return out

    def set_camera_exposures(self, exposures):
        """Replaces the per-image exposures (and resets their optimizer state). An addition, like get_camera_exposures."""
        e = np.ascontiguousarray(np.asarray(exposures, np.float32).reshape(int(self._tb._get("n_images")), 3))
        check(lib().ngpb_testbed_set_camera_exposures(self._tb._h, e.ctypes.data))


    # datasets filled from arrays while training runs (python_api.cu:806, :830-851; with Testbed.create_empty_nerf_dataset)
    n_images_for_training = property(lambda s: int(s._tb._get("n_images_for_training")), lambda s, v: s._tb._set("n_images_for_training", int(v)))

    def set_image(self, frame_idx, img, depth_img=None, depth_scale=1.0):
        """Training::set_image (python_api.cu:56-76): replaces one training image. img: (H, W, 4) float32 (or float16) linear colours with premultiplied
        alpha, as in the reference; a uint8 array is taken as sRGB RGBA8 (an addition). depth_img is accepted and unused: depth supervision is not built."""
        px, itype = _image_array(img)
        h = HostImage()
        h.pixels, h.w, h.h, h.image_type = px.ctypes.data, px.shape[1], px.shape[0], itype
        check(lib().ngpb_testbed_set_training_image(self._tb._h, int(frame_idx) & 0xFFFFFFFF, C.byref(h)))

    def set_camera_intrinsics(self, frame_idx, fx=0.0, fy=0.0, cx=-0.5, cy=-0.5, k1=0.0, k2=0.0, p1=0.0, p2=0.0):
        """Training::set_camera_intrinsics (src/testbed_nerf.cu:2502-2516): focal lengths in pixels (a non-positive one is copied from the other), principal
        point in pixels (negative: minus the fraction of the resolution), OpenCV distortion if any coefficient is non-zero."""
        check(lib().ngpb_testbed_set_camera_intrinsics(self._tb._h, int(frame_idx) & 0xFFFFFFFF, float(fx), float(fy), float(cx), float(cy), float(k1), float(k2), float(p1), float(p2)))

    # K19: importance sampling by accumulated training error (python_api.cu:817-818; train_nerf :2933-2939, :2971-3023)
    sample_focal_plane_proportional_to_error = _bool_prop("sample_focal_plane_proportional_to_error")
    sample_image_proportional_to_error = _bool_prop("sample_image_proportional_to_error")
    n_steps_between_error_map_updates = property(lambda s: int(s._tb._get("n_steps_between_error_map_updates")))
    n_steps_since_error_map_update = property(lambda s: int(s._tb._get("n_steps_since_error_map_update")))

    def get_error_map_pmf(self):
        """Sampling probabilities of the training images after the last CDF update (ErrorMap::pmf_img_cpu, testbed.h:611; the reference shows them in
        its GUI); uniform before the first update. An addition, like get_camera_exposures."""
        out = np.zeros(int(self._tb._get("n_images")), np.float32)
        check(lib().ngpb_testbed_get_error_map_pmf(self._tb._h, out.ctypes.data))
        return out


class _Nerf:
    def __init__(self, tb):
        self._tb = tb
        self.training = _Training(tb)

    cone_angle_constant = property(lambda s: s._tb._get("cone_angle_constant"), lambda s, v: s._tb._set("cone_angle_constant", float(v)))
    render_min_transmittance = property(lambda s: s._tb._get("render_min_transmittance"), lambda s, v: s._tb._set("render_min_transmittance", float(v)))
    rendering_min_transmittance = render_min_transmittance  # (both names are bound, python_api.cu:755-756)
    rgb_activation = property(lambda s: NerfActivation(int(s._tb._get("rgb_activation"))), lambda s, v: s._tb._set("rgb_activation", int(v)))
    density_activation = property(lambda s: NerfActivation(int(s._tb._get("density_activation"))), lambda s, v: s._tb._set("density_activation", int(v)))

    # scripts/run.py sets these unconditionally (run.py:106,:145). nerf.sharpen (python_api.cu:749) is the amount of sharpening the NEXT load_training_data
    # applies to the training images (Testbed::load_nerf hands it to ngp::load_nerf, src/testbed_nerf.cu:2754).
    sharpen = 0.0

    render_with_lens_distortion = False
    render_with_camera_distortion = False


def frame_buffer(height, width):
    """A float32 [height][width][4] array for a rendered frame. Page-locked when torch is importable (its caching host allocator hands the pages out without a
    page fault per 4 KB, and the library then copies the frame from the device straight into it); plain numpy memory otherwise."""
    try:
        import torch
        return torch.empty((int(height), int(width), 4), dtype=torch.float32, pin_memory=True).numpy()
    except Exception:
        return np.empty((int(height), int(width), 4), np.float32)


class Testbed:
    """pyngp.Testbed (python_api.cu:540-732). ETestbedMode::Nerf is this class; Image and Sdf return the classes of pyngp/modes.py (same constructor
    arguments); Volume is not built."""

    def __new__(cls, mode=TestbedMode.Nerf, *args, **kwargs):
        if cls is Testbed and mode in (TestbedMode.Image, TestbedMode.Sdf):
            from . import modes
            obj = object.__new__(modes.ImageTestbed if mode == TestbedMode.Image else modes.SdfTestbed)
            obj.__init__(mode, *args, **kwargs)  # (not a Testbed subclass, so Python does not call it)
            return obj
        return object.__new__(cls)

    def __init__(self, mode=TestbedMode.Nerf, data_path=None, network_config=None, device=0):
        """Testbed(mode), Testbed(mode, data_path, network_config_path) and Testbed(mode, data_path, network_config_json) (python_api.cu:542-544)."""
        if isinstance(data_path, int) and network_config is None:  # Testbed(mode, device) of earlier drafts of this module
            data_path, device = None, data_path
        if mode != TestbedMode.Nerf:
            raise RuntimeError("TestbedMode.Volume is not implemented by this build (Nerf, Image and Sdf are)")
        self.mode = mode
        self._h = C.c_void_p()
        check(lib().ngpb_testbed_create(C.byref(self._h), int(device)))
        self.nerf = _Nerf(self)
        self.training_batch_size = 1 << 18  # testbed.h:909
        self.network_config = json.loads(json.dumps(BASE_NETWORK_CONFIG))
        self._seed = 1337
        self._device = int(device)
        self._keep = None
        if data_path is not None:
            self.load_training_data(data_path)
            if isinstance(network_config, dict):
                self.reload_network_from_json(network_config)
            elif network_config is not None:
                self.reload_network_from_file(network_config)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib().ngpb_testbed_destroy(h)
            self._h = None

    # -- helpers
    def _get(self, name):
        return float(lib().ngpb_testbed_get_option(self._h, name.encode()))

    def _set(self, name, value):
        check(lib().ngpb_testbed_set_option(self._h, name.encode(), float(value)))

    # -- data
    def load_training_data(self, path):
        """Testbed::load_training_data (src/testbed.cu:97) -> load_nerf (src/testbed_nerf.cu:2735-2758): a transforms json, a directory of them, or a
        snapshot (.msgpack), which is loaded with training switched off."""
        if str(path).lower().endswith(".msgpack"):
            self.load_snapshot(path)
            self.shall_train = False
            return
        d = load_transforms(path, float(self.nerf.sharpen))
        self._dataset_scale, self._dataset_offset = d["scale"], tuple(d["offset"])
        self._from_mitsuba = bool(d.get("from_mitsuba", False))
        self.load_training_images(d["images"], d["xforms"], d["fx"], d["fy"], d["cx"], d["cy"], d["aabb_scale"], d["lenses"])
        if d.get("is_hdr"):  # load_nerf_post (src/testbed_nerf.cu:2644): HDR datasets train an exponential colour activation
            self._set("rgb_activation", int(NerfActivation.Exponential))

    def create_empty_nerf_dataset(self, n_images, aabb_scale=1, is_hdr=False):
        """Testbed::create_empty_nerf_dataset (python_api.cu:545, src/testbed_nerf.cu:2635-2641): n_images empty slots to be filled with
        nerf.training.set_image / set_camera_intrinsics / set_camera_extrinsics; nerf.training.n_images_for_training starts at 0 (train() is then a no-op)."""
        self._dataset_scale, self._dataset_offset = 1.0, (0.0, 0.0, 0.0)  # create_empty_nerf_dataset: NERF_SCALE (1.0 in this fork) and a zero offset (nerf_loader.cu:185-186)
        self._from_mitsuba = False
        check(lib().ngpb_testbed_create_empty_dataset(self._h, int(n_images), int(aabb_scale)))
        self._intrinsics = None
        self._set("rgb_activation", int(NerfActivation.Exponential if is_hdr else NerfActivation.Logistic))  # load_nerf_post (:2644)

    def load_training_images(self, images, xforms, fx, fy, cx=0.5, cy=0.5, aabb_scale=1, lenses=None):
        """Already-decoded data: images [n][h][w][4] (host) -- uint8 sRGB with straight alpha, or float16 / float32 linear colours with premultiplied alpha
        (EImageDataType Byte / Half / Float) --, xforms float32 [n][3][4] in ngp convention; lenses: per image (LensMode, 7 params) or None."""
        n = len(images)
        arr = (HostImage * n)()
        keep, intrinsics = [], []
        for i in range(n):
            px, itype = _image_array(images[i])
            keep.append(px)
            arr[i].pixels = px.ctypes.data
            arr[i].image_type = itype
            arr[i].h, arr[i].w = px.shape[0], px.shape[1]
            pick = lambda v: float(v[i]) if isinstance(v, (list, tuple, np.ndarray)) else float(v)  # scalar = the same for every image
            arr[i].fx, arr[i].fy, arr[i].cx, arr[i].cy = pick(fx), pick(fy), pick(cx), pick(cy)
            intrinsics.append((pick(fx), pick(fy), px.shape[1], px.shape[0]))
            if lenses is not None:
                arr[i].lens_mode = int(lenses[i][0])
                for k in range(7):
                    arr[i].lens_params[k] = float(lenses[i][1][k])
            cm = np.asarray(xforms[i], dtype=np.float32).reshape(3, 4).T.reshape(-1)
            for k in range(12):
                arr[i].xform[k] = float(cm[k])
        check(lib().ngpb_testbed_load_training_data(self._h, n, arr, int(aabb_scale)))
        self._intrinsics = intrinsics  # (for set_camera_to_training_view)

    def reload_network_from_file(self, path=None):
        """Testbed::reload_network_from_file (src/testbed.cu:147). Only the base NeRF architecture is built; optimizer values are honoured."""
        cfg = load_network_config(path) if path else json.loads(json.dumps(BASE_NETWORK_CONFIG))
        self.reload_network_from_json(cfg)

    def _apply_network_config(self, cfg):
        """Pushes the config values the kernels read into the testbed: table size, and the hyper-parameters of Ema(ExponentialDecay(Adam))
        (Optimizer::update_hyperparams: ema.h, exponential_decay.h, adam.h). Call before the network is (re)built."""
        self._set("log2_hashmap_size", cfg.get("encoding", {}).get("log2_hashmap_size", 19))
        ema = cfg.get("optimizer", {})
        decay = ema.get("nested", {})
        adam = decay.get("nested", {})
        base = BASE_NETWORK_CONFIG["optimizer"]
        self._set("ema_decay", ema.get("decay", base["decay"]))
        for k in ("decay_start", "decay_interval", "decay_base"):
            self._set(k, decay.get(k, base["nested"][k]))
        for k, name in (("learning_rate", "learning_rate"), ("beta1", "adam_beta1"), ("beta2", "adam_beta2"), ("epsilon", "adam_epsilon"), ("l2_reg", "adam_l2_reg")):
            self._set(name, adam.get(k, base["nested"]["nested"][k]))

    def reload_network_from_json(self, cfg, config_base_path=""):
        _validate_network_config(cfg)
        self.network_config = cfg
        self._apply_network_config(cfg)
        check(lib().ngpb_testbed_reset_network(self._h, self._seed))
        loss = str(cfg.get("loss", {}).get("otype", "Huber")).lower()
        names = {"l2": 0, "l1": 1, "mape": 2, "smape": 3, "huber": 4, "logl1": 5, "relativel2": 6}
        if loss not in names:
            raise RuntimeError(f"unknown loss type {loss}")
        self._set("loss_type", names[loss])

    def reset(self, seed=1337):
        self._seed = seed
        check(lib().ngpb_testbed_reset_network(self._h, seed))

    # -- training
    def train(self, batch_size=None):
        """Testbed::train(batch_size): exactly one optimizer step (python_api.cu:594)."""
        check(lib().ngpb_testbed_train(self._h, int(batch_size or self.training_batch_size)))

    def train_n(self, n_steps, batch_size=None):
        check(lib().ngpb_testbed_train_n(self._h, int(batch_size or self.training_batch_size), int(n_steps)))

    def frame(self):
        """Testbed::frame (src/testbed.cu:2044): one training step if shall_train; headless, no render. Returns False when training stopped."""
        if self.shall_train:
            self.train(self.training_batch_size)
        return self.shall_train

    def want_repl(self):
        """Testbed::want_repl (python_api.cu): a GUI key binding; headless sessions never ask for one."""
        return False

    tonemap_curve = property(lambda s: TonemapCurve(int(s._get("tonemap_curve"))), lambda s, v: s._set("tonemap_curve", int(TonemapCurve(int(v)))))  # m_tonemap_curve (testbed.h:847)

    shall_train = property(lambda s: bool(s._get("shall_train")), lambda s, v: s._set("shall_train", 1.0 if v else 0.0))
    training_step = property(lambda s: int(lib().ngpb_testbed_training_step(s._h)))
    loss = property(lambda s: float(lib().ngpb_testbed_loss(s._h)))
    n_params = property(lambda s: _CountInt(lib().ngpb_testbed_n_params(s._h)))  # Testbed::n_params() (python_api.cu:612)
    n_encoding_params = property(lambda s: _CountInt(max(int(lib().ngpb_testbed_n_params(s._h)) - 10240, 0)))  # Testbed::n_encoding_params() (:613): the hash-grid entries
    # options of the reference's Testbed outside this path (python_api.cu:656-694): defaults only
    shall_train_encoding = _unbuilt("testbed", "shall_train_encoding", True)
    shall_train_network = _unbuilt("testbed", "shall_train_network", True)
    max_level_rand_training = _unbuilt("testbed", "max_level_rand_training", False)
    render_with_rolling_shutter = _unbuilt("testbed", "render_with_rolling_shutter", False)
    dlss = _unbuilt("testbed", "dlss", False)
    dynamic_res = _unbuilt("testbed", "dynamic_res", False)
    # the fork's mask / camera-model members act on the Blender path through RenderRequest (request_nerf_render_sync); Testbed.render is the classic path
    render_masks = _unbuilt("testbed", "render_masks", [])

    def first_training_view(self):
        """Testbed::first_training_view (src/testbed_nerf.cu): the camera moves to training image 0."""
        self.set_camera_to_training_view(0)

    def set_camera_to_training_view(self, trainview):
        """Testbed::set_camera_to_training_view: camera = the image's current training transform, focal length = the image's (relative to the
        resolution along fov_axis). The classic render keeps the principal point at the centre (an off-centre one is not built)."""
        out = np.zeros(12, np.float32)
        check(lib().ngpb_testbed_get_camera_extrinsics(self._h, int(trainview), out.ctypes.data, None, None))
        self.camera_matrix = out.reshape(4, 3).T.copy()
        intr = self.__dict__.get("_intrinsics")
        if intr is not None and 0 <= int(trainview) < len(intr):
            fx, fy, w, h = intr[int(trainview)]
            res = (w, h)[self.fov_axis]
            self._relative_focal_length = (fx / res, fy / res)

    @property
    def background_color(self):  # m_background_color, RGBA (testbed.h:875)
        return [self._get("background_color_" + c) for c in "rgba"]

    @background_color.setter
    def background_color(self, v):
        for c, x in zip("rgba", v):
            self._set("background_color_" + c, float(x))

    color_space = property(lambda s: ColorSpace(int(s._get("color_space"))), lambda s, v: s._set("color_space", int(v)))

    def stats(self):
        s = (C.c_uint64 * 4)()
        check(lib().ngpb_testbed_stats(self._h, s))
        return dict(rays_per_batch=int(s[0]), measured_batch_size_before_compaction=int(s[1]), measured_batch_size=int(s[2]), gpu_launches=int(s[3]),
                    h2d_bytes=int(self._get("h2d_bytes")), d2h_bytes=int(self._get("d2h_bytes")))

    def init_data_parallel(self, rank, world, group=None):
        """Joins this Testbed to a data-parallel group of `world` processes (one per GPU). Needs torch.distributed to be initialised: it carries
        the 128-byte NCCL unique id from rank 0 to the others; the gradient all-reduce itself runs inside libngpb200.so on its own communicator."""
        if world == 1:
            check(lib().ngpb_testbed_init_data_parallel(self._h, 0, 1, None))
            return
        import torch
        import torch.distributed as dist
        uid = C.create_string_buffer(128)
        if rank == 0:
            check(lib().ngpb_nccl_unique_id(uid))
        t = torch.frombuffer(bytearray(uid.raw), dtype=torch.uint8).clone()
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0, group=group)
        raw = bytes(t.cpu().numpy().tobytes())
        check(lib().ngpb_testbed_init_data_parallel(self._h, int(rank), int(world), raw))

    STAGES = ["sampling", "encode_inference", "mlp_inference", "loss", "encode_train", "mlp_train", "encode_backward", "optimizer",
              "density_grid", "allreduce"]

    def profile_stages(self, on=True):
        self._set("profile_stages", 1.0 if on else 0.0)

    def stage_times(self, reset=False):
        """Per-stage device time (CUDA events on the testbed's stream): {stage: (ms, calls, units)} accumulated since the last reset."""
        n = len(self.STAGES)
        ms = (C.c_double * n)(); calls = (C.c_uint64 * n)(); units = (C.c_uint64 * n)()
        check(lib().ngpb_testbed_stage_times(self._h, ms, calls, units, int(reset)))
        return {name: (float(ms[i]), int(calls[i]), int(units[i])) for i, name in enumerate(self.STAGES)}

    @property
    def stream(self):
        """cudaStream_t (as int) every kernel of this testbed runs on."""
        return int(lib().ngpb_testbed_stream(self._h) or 0)

    # -- snapshots
    def save_snapshot(self, path, include_optimizer_state=False):
        """Testbed::save_snapshot (src/testbed.cu:3008): msgpack of the network config + snapshot object, readable by the reference."""
        import msgpack
        st = TrainingState()
        check(lib().ngpb_testbed_get_training_state(self._h, C.byref(st)))
        _, _, ema = self.get_params()
        grid, _ = self.get_density_grid()
        opt = None
        if include_optimizer_state:
            n = self.n_params
            fm = np.empty(n, np.float32); sm = np.empty(n, np.float32); ps = np.empty(n, np.uint32)
            check(lib().ngpb_testbed_get_optimizer_state(self._h, fm.ctypes.data_as(C.c_void_p), sm.ctypes.data_as(C.c_void_p), ps.ctypes.data_as(C.c_void_p)))
            opt = dict(current_step=st.optimizer_step, learning_rate=st.learning_rate, learning_rate_factor=st.learning_rate_factor, first_moments=fm, second_moments=sm, param_steps=ps)
        half = 0.5 * min(128, int(self._get("aabb_scale")))
        aabb = [0.5 - half] * 3 + [0.5 + half] * 3
        cfg = build_snapshot(self.network_config, ema, grid, int(self._get("aabb_scale")), aabb, st.training_step, st.loss, st.rays_per_batch,
                             st.measured_batch_size, st.measured_batch_size_before_compaction, opt, (self._dataset_scale, self._dataset_offset, self._from_mitsuba))
        with open(path, "wb") as f:
            f.write(msgpack.packb(cfg, use_bin_type=True))

    def load_snapshot(self, path):
        """Testbed::load_snapshot (src/testbed.cu:3044): restores network, occupancy grid and training counters; without a loaded dataset the
        session can render but not train."""
        import msgpack
        with open(path, "rb") as f:
            cfg = msgpack.unpackb(f.read(), raw=False, strict_map_key=False)
        snap = parse_snapshot(cfg)
        _validate_network_config(snap["network_config"])
        self.network_config = snap["network_config"]
        if snap["dataset_transform"] is not None:
            self._dataset_scale, self._dataset_offset = snap["dataset_transform"]
            self._from_mitsuba = bool(snap.get("dataset_from_mitsuba", False))
        self._apply_network_config(self.network_config)
        check(lib().ngpb_testbed_configure(self._h, snap["aabb_scale"], self._seed))
        params = np.ascontiguousarray(snap["params_half"], np.float16)
        check(lib().ngpb_testbed_set_params_half(self._h, params.ctypes.data_as(C.c_void_p), int(params.shape[0])))
        if snap["density_grid"].size:
            grid = np.ascontiguousarray(snap["density_grid"], np.float32)
            check(lib().ngpb_testbed_set_density_grid(self._h, grid.ctypes.data_as(C.c_void_p), int(grid.shape[0])))
        st = TrainingState()
        st.training_step = snap["training_step"]; st.loss = snap["loss"]
        st.rays_per_batch = int(snap["rgb"].get("rays_per_batch", 1 << 12))
        st.measured_batch_size = int(snap["rgb"].get("measured_batch_size", 0))
        st.measured_batch_size_before_compaction = int(snap["rgb"].get("measured_batch_size_before_compaction", 0))
        opt = snap["optimizer"]
        # without an optimizer block the reference keeps the fresh optimizer reset_network built: step 0, factor 1 (Trainer::deserialize, trainer.h:290-310)
        st.optimizer_step = opt["current_step"] if opt else 0
        st.learning_rate = opt["learning_rate"] if opt else 0.0
        st.learning_rate_factor = opt["learning_rate_factor"] if opt else 1.0
        check(lib().ngpb_testbed_set_training_state(self._h, C.byref(st)))
        if opt:
            n = self.n_params
            ps = opt["param_steps"] if opt["param_steps"] is not None else np.zeros(n, np.uint32)
            fm, sm, ps = (np.ascontiguousarray(a) for a in (opt["first_moments"], opt["second_moments"], ps))
            if not (fm.shape[0] == sm.shape[0] == ps.shape[0] == n):
                raise RuntimeError("snapshot optimizer state does not match the parameter count")
            check(lib().ngpb_testbed_set_optimizer_state(self._h, fm.ctypes.data_as(C.c_void_p), sm.ctypes.data_as(C.c_void_p), ps.ctypes.data_as(C.c_void_p)))

    # -- parameters / state
    def get_params(self):
        n = self.n_params
        w = np.empty(n, np.float32); h = np.empty(n, np.float16); e = np.empty(n, np.float16)
        check(lib().ngpb_testbed_get_params(self._h, w.ctypes.data_as(C.c_void_p), h.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p)))
        return w, h, e

    def set_params(self, w_fp32):
        w = np.ascontiguousarray(w_fp32, dtype=np.float32)
        if w.shape[0] != self.n_params:
            raise RuntimeError("set_params: wrong parameter count")
        check(lib().ngpb_testbed_set_params(self._h, w.ctypes.data_as(C.c_void_p)))

    def get_density_grid(self):
        n_casc = int(self._get("max_cascade")) + 1
        g = np.empty(128 ** 3 * n_casc, np.float32); b = np.empty(128 ** 3, np.uint8)
        check(lib().ngpb_testbed_get_density_grid(self._h, g.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)))
        return g, b

    # -- render
    def render(self, width, height, spp=1, linear=True, start_t=-1.0, end_t=-1.0, fps=30.0, shutter_fraction=1.0):
        """Testbed::render (python_api.cu:567 -> render_to_cpu :132): float32 [H][W][4], RGBA after accumulation over `spp` samples and
        tone mapping; `linear=False` converts to sRGB. Static camera only (camera paths / motion blur are outside the built scope)."""
        if start_t >= 0.0:
            raise RuntimeError("camera-path rendering (start_t >= 0) is outside the built scope")
        cam = np.asarray(self.camera_matrix, dtype=np.float32).reshape(3, 4).T.reshape(-1).copy()
        # Testbed::calc_focal_length (src/testbed.cu:2589): relative focal length x resolution[fov_axis] x zoom, same for x and y
        rel = self._relative_focal_length
        res_axis = (width, height)[self.fov_axis]
        fx, fy = rel[0] * res_axis, rel[1] * res_axis
        out = frame_buffer(height, width)
        ns = C.c_uint64(0)
        check(lib().ngpb_testbed_render(self._h, cam.ctypes.data_as(C.c_void_p), int(width), int(height), C.c_float(fx), C.c_float(fy), int(spp), int(bool(linear)),
                                        out.ctypes.data_as(C.c_void_p), C.byref(ns)))
        self.last_render_samples = int(ns.value)
        self.last_render_ms = float(lib().ngpb_testbed_last_render_ms(self._h))
        return out

    # -- the Blender multi-NeRF renderer (python_api.cu:214-262, :577-585)
    def _fields_for(self, descriptors):
        """RenderData::update_nerfs (nerf/render_data.cuh:44-80): fields are cached by snapshot path, dropped when no descriptor names them."""
        cache = self.__dict__.setdefault("_bl_fields", {})
        wanted = {d.snapshot_path for d in descriptors}
        for k in [k for k in cache if k not in wanted]:
            del cache[k]
        for d in descriptors:
            if d.snapshot_path not in cache:
                cache[d.snapshot_path] = Field.from_snapshot(d.snapshot_path, self._device)
        return [cache[d.snapshot_path] for d in descriptors]

    def request_nerf_render_sync(self, render_request):
        """Testbed::bl_request_nerf_render_sync (python_api.cu:233): float32 [H][W][4] of all the request's NeRFs composited front to back."""
        out_p, cam_p = render_request.output, render_request.camera
        w, h = out_p.resolution
        result = frame_buffer(h, w)
        if self.__dict__.get("_currently_rendering", False):
            result.fill(0.0)
            return result
        self._currently_rendering = True
        try:
            fields = self._fields_for(render_request.nerfs)
            rq = BlenderRequest()
            rq.width, rq.height, rq.mip, rq.flip_y = w, h, out_p.ds.mip, int(out_p.flip_y)
            cm = cam_p.transform.T.reshape(-1)
            for k in range(12):
                rq.camera[k] = float(cm[k])
            rq.focal_length, rq.near_distance, rq.color_space, rq.exposure = cam_p.focal_length, cam_p.near_distance, int(out_p.color_space), out_p.exposure
            for k in range(4):
                rq.background_color[k] = out_p.background_color[k]
            rq.camera_model, rq.aperture_size, rq.focus_z, rq.tonemap_curve = int(cam_p.model), cam_p.aperture_size, cam_p.focus_z, int(out_p.tonemap_curve)
            sq, qh = cam_p.spherical_quadrilateral, cam_p.quadrilateral_hexahedron
            if sq is not None:
                rq.spherical_quadrilateral[0], rq.spherical_quadrilateral[1], rq.spherical_quadrilateral[2] = sq.width, sq.height, sq.curvature
            if qh is not None:
                flat = np.concatenate([q for quad in (qh.front, qh.back) for q in (quad.tl, quad.tr, quad.bl, quad.br)])
                for k in range(24):
                    rq.quadrilateral_hexahedron[k] = float(flat[k])
            keep = [_mask_array(render_request.modifiers.masks if render_request.modifiers is not None else [])]
            if keep[0] is not None:
                rq.n_masks, rq.masks = len(keep[0]), keep[0]
            inst = (NerfInstance * max(1, len(fields)))()
            for i, (d, f) in enumerate(zip(render_request.nerfs, fields)):
                inst[i].field = f._h
                for k in range(3):
                    inst[i].aabb[k], inst[i].aabb[3 + k] = float(d.aabb.min[k]), float(d.aabb.max[k])
                t = d.transform.T.reshape(-1)
                for k in range(16):
                    inst[i].transform[k] = float(t[k])
                inst[i].opacity = d.opacity
                arr = _mask_array(d.modifiers.masks if d.modifiers is not None else [])
                if arr is not None:
                    keep.append(arr)
                    inst[i].n_masks, inst[i].masks = len(arr), arr
            ns, nl = C.c_uint64(0), C.c_uint32(0)
            check(lib().ngpb_blender_render(C.c_void_p(self.stream or None), C.byref(rq), len(fields), inst, result.ctypes.data_as(C.c_void_p), C.byref(ns), C.byref(nl)))
            self.last_render_samples, self.last_render_launches = int(ns.value), int(nl.value)
        finally:
            self._currently_rendering = False
        return result

    def request_nerf_render_async(self, render_request, render_callback):
        """Testbed::bl_request_nerf_render_async (python_api.cu:214): renders on a detached thread and hands the image to the callback;
        a request that arrives while one is in flight is dropped, as in the reference."""
        import threading
        if self.__dict__.get("_currently_rendering", False):
            return
        t = threading.Thread(target=lambda: render_callback(self.request_nerf_render_sync(render_request)), daemon=True)
        t.start()
        self._render_thread = t

    fov_axis = 1  # m_fov_axis (testbed.h:528)
    _relative_focal_length = (1.0, 1.0)  # m_relative_focal_length (testbed.h:527)
    camera_matrix = np.eye(4, dtype=np.float32)[:3]  # m_camera, 3x4, ngp convention
    _dataset_scale, _dataset_offset = 1.0, (0.0, 0.0, 0.0)
    _from_mitsuba = False  # NerfDataset::from_mitsuba (nerf_loader.h:99): set by a transforms.json with "normal_mts_args"

    @property
    def fov(self):  # degrees along fov_axis (Testbed::fov, src/testbed.cu:2153)
        return math.degrees(2.0 * math.atan(0.5 / self._relative_focal_length[self.fov_axis]))

    @fov.setter
    def fov(self, deg):  # Testbed::set_fov (:2157)
        f = 0.5 / math.tan(0.5 * math.radians(deg))
        self._relative_focal_length = (f, f)

    snap_to_pixel_centers = property(lambda s: bool(s._get("render_snap_to_pixel_centers")), lambda s, v: s._set("render_snap_to_pixel_centers", 1.0 if v else 0.0))
    exposure = property(lambda s: s._get("exposure"), lambda s, v: s._set("exposure", float(v)))
    render_near_distance = property(lambda s: s._get("render_near_distance"), lambda s, v: s._set("render_near_distance", float(v)))

    def set_nerf_camera_matrix(self, m):
        """python_api.cu:681 -> Testbed::set_nerf_camera_matrix: a NeRF-convention camera-to-world matrix, converted with the dataset's
        scale / offset (NerfDataset::nerf_matrix_to_ngp, nerf_loader.h:113-132)."""
        self.camera_matrix = nerf_matrix_to_ngp(np.asarray(m, dtype=np.float32), self._dataset_scale, self._dataset_offset, self._from_mitsuba)
