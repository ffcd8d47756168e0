"""pyngp.Testbed for ETestbedMode::Image and ETestbedMode::Sdf (python_api.cu:540-732) on top of the C ABI's ngpb_model: the same hash-grid and fully fused
MLP kernels as the NeRF path behind the reference's names. Image: load_training_data (8-bit image files, the .bin half format), train / frame, render,
compute_image_mse, snapshots. SDF: training on supplied (point, distance) pairs -- override_sdf_training_data (python_api.cu:74-104) -- after
load_training_data read the mesh's bounding box; BVH distance queries, sphere tracing and marching cubes are outside the built path."""
import ctypes as C
import json
import math
import os

import numpy as np


def _pyngp():
    import pyngp
    return pyngp


IMAGE_NETWORK_CONFIG = {  # configs/image/base.json
    "loss": {"otype": "L2"},
    "optimizer": {"otype": "ExponentialDecay", "decay_start": 20000, "decay_interval": 10000, "decay_base": 0.33,
                  "nested": {"otype": "Adam", "learning_rate": 1e-2, "beta1": 0.9, "beta2": 0.99, "epsilon": 1e-15, "l2_reg": 1e-6}},
    "encoding": {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 24, "base_resolution": 16},
    "network": {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 2},
}
SDF_NETWORK_CONFIG = {  # configs/sdf/base.json
    "loss": {"otype": "MAPE"},
    "optimizer": {"otype": "Ema", "decay": 0.95, "nested": {"otype": "ExponentialDecay", "decay_start": 10000, "decay_interval": 5000, "decay_base": 0.33,
                  "nested": {"otype": "Adam", "learning_rate": 1e-4, "beta1": 0.9, "beta2": 0.99, "epsilon": 1e-15, "l2_reg": 1e-6}}},
    "encoding": {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 19, "base_resolution": 16},
    "network": {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 2},
}
_LOSSES = {"l2": 0, "mape": 1, "relativel2": 2}


def model_config_struct(cfg, n_pos_dims, n_output_dims, desired_resolution, seed):
    """Network config json -> ngpb_model_config, with Testbed::reset_network's derived values (src/testbed.cu:2290-2325): base_resolution defaults to
    2^(log2_hashmap_size / n_pos), per_level_scale = exp(ln(desired_resolution / base) / (L - 1)) unless the config gives one."""
    P = _pyngp()
    enc, net = cfg.get("encoding", {}), cfg.get("network", {})
    def req(ok, what):
        if not ok:
            raise RuntimeError(f"unsupported network config: {what} (this build implements HashGrid 16 x 2 + FullyFusedMLP 64 x 2 hidden layers, ReLU)")
    req("grid" in str(enc.get("otype", "HashGrid")).lower(), "encoding.otype")
    req(enc.get("n_levels", 16) == 16 and enc.get("n_features_per_level", 2) == 2, "encoding levels / features")
    req(str(net.get("otype", "FullyFusedMLP")).lower() == "fullyfusedmlp" and net.get("n_neurons", 64) == 64 and net.get("n_hidden_layers", 2) == 2, "network")
    req(str(net.get("activation", "ReLU")).lower() == "relu" and str(net.get("output_activation", "None")).lower() == "none", "network activations")
    loss = str(cfg.get("loss", {}).get("otype", "L2")).lower()
    req(loss in _LOSSES, f"loss.otype {loss}")
    log2_t = int(enc.get("log2_hashmap_size", 15))
    base = int(enc.get("base_resolution", 0)) or (1 << (log2_t // n_pos_dims))
    pls = float(enc.get("per_level_scale", 0.0))  # <= 0: derived inside the library (host libm, like the reference)
    # optimizer: [Ema(] ExponentialDecay( Adam ) [)]
    o = cfg.get("optimizer", {})
    use_ema = str(o.get("otype", "")).lower() == "ema"
    decay = o.get("nested", {}) if use_ema else o
    adam = decay.get("nested", {}) if str(decay.get("otype", "")).lower() == "exponentialdecay" else decay
    req(str(adam.get("otype", "Adam")).lower() == "adam", "optimizer (Adam, optionally inside ExponentialDecay and Ema)")
    opt = P.Optimizer()
    P.lib().ngpb_optimizer_init(C.byref(opt))
    opt.learning_rate = float(adam.get("learning_rate", 1e-3)); opt.beta1 = float(adam.get("beta1", 0.9)); opt.beta2 = float(adam.get("beta2", 0.999))
    opt.epsilon = float(adam.get("epsilon", 1e-8)); opt.l2_reg = float(adam.get("l2_reg", 1e-8))
    opt.ema_decay = float(o.get("decay", 0.99)) if use_ema else 0.0
    if adam is not decay:
        opt.decay_start = int(decay.get("decay_start", 10000)); opt.decay_interval = int(decay.get("decay_interval", 10000)); opt.decay_base = float(decay.get("decay_base", 0.33))
    else:
        opt.decay_start = 0xFFFFFFFF; opt.decay_interval = 0; opt.decay_base = 1.0
    mc = P.ModelConfig()
    mc.n_pos_dims, mc.n_output_dims, mc.n_levels, mc.log2_hashmap_size, mc.base_resolution = n_pos_dims, n_output_dims, 16, log2_t, base
    mc.per_level_scale, mc.desired_resolution = pls, float(desired_resolution)
    mc.loss, mc.use_ema, mc.optimizer, mc.seed = _LOSSES[loss], int(use_ema), opt, int(seed)
    return mc


class _ImageTraining:
    def __init__(self, tb):
        self._tb = tb

    snap_to_pixel_centers = property(lambda s: s._tb._snap, lambda s, v: s._tb._set_model_option("snap_to_pixel_centers", bool(v)))
    linear_colors = property(lambda s: s._tb._linear, lambda s, v: s._tb._set_model_option("linear_colors", bool(v)))


class _ImageNs:
    def __init__(self, tb):
        self.training = _ImageTraining(tb)
        self.random_mode = 0  # ERandomMode::Stratified is the only one built


class ModelTestbedBase:
    """What Image and Sdf share: the model handle, train / frame / loss / training_step, parameters, snapshots."""
    _default_config = None
    _n_pos, _n_out = 0, 0

    def _init_common(self, mode, device):
        P = _pyngp()
        self.mode = mode
        self._m = C.c_void_p()
        self._device = int(device)
        self._seed = 1337
        self.network_config = json.loads(json.dumps(self._default_config))
        self.training_batch_size = 1 << 18
        self.shall_train = True
        self._data_loaded = False
        self._snap, self._linear = True, False
        self.background_color = [0.0, 0.0, 0.0, 1.0]
        self.exposure = 0.0
        self.color_space = P.ColorSpace.SRGB
        self.tonemap_curve = P.TonemapCurve.Identity
        self.snap_to_pixel_centers = False

    def __del__(self):
        m = getattr(self, "_m", None)
        if m:
            try:
                _pyngp().lib().ngpb_model_destroy(m)
            except ImportError:  # interpreter shutdown
                pass
            self._m = None

    def _desired_resolution(self):
        return 2048.0

    def _build(self):
        P = _pyngp()
        if self._m:
            P.lib().ngpb_model_destroy(self._m)
            self._m = C.c_void_p()
        mc = model_config_struct(self.network_config, self._n_pos, self._n_out, self._desired_resolution(), self._seed)
        P.check(P.lib().ngpb_model_create(C.byref(self._m), self._device, C.byref(mc)))
        self._model_config = mc
        self._set_model_option("snap_to_pixel_centers", self._snap)
        self._set_model_option("linear_colors", self._linear)

    def _set_model_option(self, name, value):
        if name == "snap_to_pixel_centers":
            self._snap = bool(value)
        if name == "linear_colors":
            self._linear = bool(value)
        if self._m:
            P = _pyngp()
            P.check(P.lib().ngpb_model_set_option(self._m, name.encode(), float(value)))

    def _need_model(self):
        if not self._m:
            raise RuntimeError("no training data loaded")

    # -- network
    def reload_network_from_file(self, path=None):
        cfg = _pyngp().load_network_config(path) if path else json.loads(json.dumps(self._default_config))
        self.reload_network_from_json(cfg)

    def reload_network_from_json(self, cfg, config_base_path=""):
        self.network_config = cfg
        if self._data_loaded:
            self._build()
            self._upload_data()

    def reset(self, seed=1337):
        self._seed = seed
        if self._m:
            P = _pyngp()
            P.check(P.lib().ngpb_model_reset(self._m, int(seed)))

    # -- training
    def train(self, batch_size=None):
        """Testbed::train(batch_size) (src/testbed.cu:2480-2560): one optimizer step; the loss scalar is read every 16th step."""
        self._need_model()
        self._train_once(int(batch_size or self.training_batch_size), self.training_step % 16 == 0)

    def train_n(self, n_steps, batch_size=None):
        for _ in range(int(n_steps)):
            self.train(batch_size)

    def frame(self):
        if self.shall_train:
            self.train(self.training_batch_size)
        return self.shall_train

    def want_repl(self):
        return False

    training_step = property(lambda s: int(_pyngp().lib().ngpb_model_training_step(s._m)) if s._m else 0)
    loss = property(lambda s: float(_pyngp().lib().ngpb_model_loss(s._m)) if s._m else 0.0)
    n_params = property(lambda s: int(_pyngp().lib().ngpb_model_n_params(s._m)) if s._m else 0)

    def get_params(self):
        P = _pyngp()
        self._need_model()
        n = self.n_params
        w = np.empty(n, np.float32); h = np.empty(n, np.float16); e = np.empty(n, np.float16)
        P.check(P.lib().ngpb_model_get_params(self._m, w.ctypes.data_as(C.c_void_p), h.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p)))
        return w, h, e

    def set_params_half(self, params):
        P = _pyngp()
        self._need_model()
        params = np.ascontiguousarray(params, np.float16)
        P.check(P.lib().ngpb_model_set_params_half(self._m, params.ctypes.data_as(C.c_void_p), int(params.shape[0])))

    def inference(self, positions, use_inference_params=True):
        """m_network->inference on host positions [n][n_pos] -> [n][n_out] float32 (test / tooling helper; torch moves the buffers)."""
        import torch
        P = _pyngp()
        self._need_model()
        pos = torch.from_numpy(np.ascontiguousarray(positions, np.float32)).cuda(self._device)
        out = torch.zeros((pos.shape[0], self._n_out), dtype=torch.float32, device=pos.device)
        torch.cuda.synchronize()
        P.check(P.lib().ngpb_model_inference(self._m, C.c_void_p(pos.data_ptr()), int(pos.shape[0]), C.c_void_p(out.data_ptr()), int(bool(use_inference_params))))
        torch.cuda.synchronize()  # (the model works on its own stream; ngpb_model_inference leaves the result in flight)
        P.check(P.lib().ngpb_model_get_training_batch(self._m, 0, None, None))  # synchronises the model's stream
        return out.cpu().numpy()

    def training_batch(self, n):
        """The batch the last training step used: (positions [n][n_pos], targets [n][n_out])."""
        P = _pyngp()
        pos = np.empty((n, self._n_pos), np.float32); tgt = np.empty((n, self._n_out), np.float32)
        P.check(P.lib().ngpb_model_get_training_batch(self._m, int(n), pos.ctypes.data_as(C.c_void_p), tgt.ctypes.data_as(C.c_void_p)))
        return pos, tgt

    # -- snapshots (Testbed::save_snapshot / load_snapshot, src/testbed.cu:3008-3106): network config + parameters + counters
    def save_snapshot(self, path, include_optimizer_state=False):
        import msgpack
        if include_optimizer_state:
            raise RuntimeError("optimizer state in image / SDF snapshots is outside the built scope")
        P = _pyngp()
        _, _, inference = self.get_params()
        cfg = json.loads(json.dumps(self.network_config))
        cfg["snapshot"] = {"n_params": int(inference.shape[0]), "params_type": "__half", "params_binary": inference.tobytes(), "version": P.SNAPSHOT_FORMAT_VERSION,
                           "density_grid_size": 128, "density_grid_binary": b"", "training_step": self.training_step, "loss": self.loss,
                           "aabb": {"min": [0.0] * 3, "max": [1.0] * 3}, "bounding_radius": 1.0, "nerf": {"aabb_scale": 1, "rgb": {}}}
        with open(path, "wb") as f:
            f.write(msgpack.packb(cfg, use_bin_type=True))

    def load_snapshot(self, path):
        import msgpack
        P = _pyngp()
        with open(path, "rb") as f:
            cfg = msgpack.unpackb(f.read(), raw=False, strict_map_key=False)
        snap = P.parse_snapshot(cfg)
        if not self._data_loaded:
            raise RuntimeError("load_snapshot in image / SDF mode needs the training data first (the grid geometry depends on it)")
        self.network_config = snap["network_config"]
        self._build()
        self._upload_data()
        self.set_params_half(snap["params_half"])
        P.check(P.lib().ngpb_model_set_training_step(self._m, int(snap["training_step"])))


class ImageTestbed(ModelTestbedBase):
    """pyngp.Testbed(TestbedMode.Image)."""
    _default_config = IMAGE_NETWORK_CONFIG
    _n_pos, _n_out = 2, 3

    def __init__(self, mode=None, data_path=None, network_config=None, device=0):
        P = _pyngp()
        self._init_common(P.TestbedMode.Image, device)
        self.image = _ImageNs(self)
        self._pixels = None
        self.scale, self.image_pos, self.screen_center = 1.0, (0.0, 0.0), (0.5, 0.5)  # m_scale, m_image.pos, m_screen_center after reset_camera
        if data_path is not None:
            self.load_training_data(data_path)
            if isinstance(network_config, dict):
                self.reload_network_from_json(network_config)
            elif network_config is not None:
                self.reload_network_from_file(network_config)

    def _desired_resolution(self):
        return max(self._res) / 2.0  # m_image.resolution.maxCoeff() / 2 (src/testbed.cu:2311)

    def load_training_data(self, path):
        """Testbed::load_image (src/testbed_image.cu:349-432): `.bin` = {int32 height, int32 width, RGBA half pixels}; `.exr` = float RGBA as stored; anything else is an
        8-bit file decoded to RGBA8 and converted like load_stbi (sRGB -> linear, premultiplied by alpha)."""
        path = str(path)
        if not os.path.exists(path):
            raise RuntimeError(f"{path} does not exist.")
        ext = os.path.splitext(path)[1].lower()
        if ext == ".exr":  # load_exr_image (src/testbed_image.cu:385-398): the file's floats as they are (EDataType::Float)
            self.load_image_data(_pyngp().load_exr_float(path))
        elif ext == ".bin":
            with open(path, "rb") as f:
                h, w = np.frombuffer(f.read(8), np.int32)
                px = np.frombuffer(f.read(int(h) * int(w) * 8), np.float16).reshape(int(h), int(w), 4)
            self.load_image_data(px)
        else:
            from PIL import Image as PILImage
            self.load_image_data(np.asarray(PILImage.open(path).convert("RGBA"), dtype=np.uint8))

    def load_image_data(self, pixels):
        """Already-decoded image [h][w][4]: uint8 (an 8-bit file's pixels), float16 (.bin) or float32 (linear, premultiplied: what load_stbi / load_exr leave in m_image.data)."""
        px = np.ascontiguousarray(pixels)
        if px.ndim != 3 or px.shape[2] != 4 or px.dtype not in (np.uint8, np.float16, np.float32):
            raise RuntimeError("image should be (H,W,C) where C=4, uint8 / float16 / float32")
        self._pixels = px
        self._res = (px.shape[1], px.shape[0])
        self._data_loaded = True
        self._build()
        self._upload_data()

    def _upload_data(self):
        P = _pyngp()
        px = self._pixels
        if px.dtype == np.uint8:
            P.check(P.lib().ngpb_model_set_image_rgba8(self._m, px.ctypes.data_as(C.c_void_p), px.shape[1], px.shape[0]))
        else:
            P.check(P.lib().ngpb_model_set_image(self._m, px.ctypes.data_as(C.c_void_p), px.shape[1], px.shape[0], int(px.dtype == np.float16)))

    def _train_once(self, batch, get_loss):
        P = _pyngp()
        P.check(P.lib().ngpb_model_train_image(self._m, batch, int(get_loss)))

    def compute_image_mse(self, quantize_to_byte=False):
        P = _pyngp()
        self._need_model()
        out = C.c_float()
        P.check(P.lib().ngpb_model_image_mse(self._m, int(bool(quantize_to_byte)), C.byref(out)))
        return float(out.value)

    def render(self, width=1920, height=1080, spp=1, linear=True, start_t=-1.0, end_t=-1.0, fps=30.0, shutter_fraction=1.0):
        """Testbed::render_to_cpu -> render_frame -> render_image (src/testbed_image.cu:285-347) + accumulate + tonemap."""
        P = _pyngp()
        self._need_model()
        if end_t >= 0 or start_t >= 0:
            raise RuntimeError("camera paths are outside the built scope")
        out = np.empty((int(height), int(width), 4), np.float32)
        view = (C.c_float * 5)(self.scale, self.image_pos[0], self.image_pos[1], self.screen_center[0], self.screen_center[1])
        bg = (C.c_float * 4)(*[float(v) for v in self.background_color])
        P.check(P.lib().ngpb_model_render_image(self._m, int(width), int(height), int(spp), view, int(bool(self.snap_to_pixel_centers)), int(self.color_space), int(not linear),
                                                C.c_float(self.exposure), bg, int(self.tonemap_curve), out.ctypes.data_as(C.c_void_p)))
        return out


class _SdfNs:
    def __init__(self):
        self.mesh_scale = 1.0


class SdfTestbed(ModelTestbedBase):
    """pyngp.Testbed(TestbedMode.Sdf), training on supplied pairs."""
    _default_config = SDF_NETWORK_CONFIG
    _n_pos, _n_out = 3, 1

    def __init__(self, mode=None, data_path=None, network_config=None, device=0):
        P = _pyngp()
        self._init_common(P.TestbedMode.Sdf, device)
        self.sdf = _SdfNs()
        self._raw_aabb = (np.zeros(3, np.float32), np.ones(3, np.float32))
        self._pairs = None
        self._data_loaded = True  # the grid geometry does not depend on the data
        self._build()
        if data_path is not None:
            self.load_training_data(data_path)
            if isinstance(network_config, dict):
                self.reload_network_from_json(network_config)
            elif network_config is not None:
                self.reload_network_from_file(network_config)

    def load_training_data(self, path):
        """Testbed::load_mesh (src/testbed_sdf.cu:989-1064) as far as supplied-pair training needs it: the ascii .obj's vertices give m_raw_aabb (inflated
        by 0.5 % of its diagonal) and mesh_scale = its largest extent, which override_sdf_training_data uses to map points into the unit cube."""
        path = str(path)
        if not path.lower().endswith(".obj"):
            raise RuntimeError("Sdf data path must be a mesh in ascii .obj format (binary .stl is outside the built scope).")
        verts = []
        with open(path) as f:
            for line in f:
                if line.startswith("v "):
                    verts.append([float(x) for x in line.split()[1:4]])
        if not verts:
            raise RuntimeError("mesh has no vertices")
        v = np.asarray(verts, np.float32)
        lo, hi = v.min(0), v.max(0)
        amount = np.float32(np.linalg.norm((hi - lo).astype(np.float32)) * np.float32(0.005))
        lo, hi = lo - amount, hi + amount
        self._raw_aabb = (lo.astype(np.float32), hi.astype(np.float32))
        self.sdf.mesh_scale = float((hi - lo).max())

    def override_sdf_training_data(self, points, distances):
        """python_api.cu:74-104: points [n][3] in mesh coordinates and distances [n] become the training pool, mapped into the unit cube."""
        pts = np.ascontiguousarray(points, np.float32); d = np.ascontiguousarray(distances, np.float32)
        if pts.ndim != 2 or d.ndim != 1 or pts.shape[0] != d.shape[0] or pts.shape[1] != 3:
            print("Invalid Points<->Distances data")  # the reference logs and returns
            return
        lo, hi = self._raw_aabb
        s = np.float32(self.sdf.mesh_scale)
        pos = (pts - lo) / s + np.float32(0.5) * (np.ones(3, np.float32) - (hi - lo) / s)
        self.set_unit_cube_pairs(pos.astype(np.float32), (d / s).astype(np.float32))

    def set_unit_cube_pairs(self, positions, distances):
        self._pairs = (np.ascontiguousarray(positions, np.float32), np.ascontiguousarray(distances, np.float32))
        self._upload_data()

    def _upload_data(self):
        if self._pairs is None:
            return
        P = _pyngp()
        pos, d = self._pairs
        P.check(P.lib().ngpb_model_set_sdf_data(self._m, pos.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), int(d.shape[0])))

    def _train_once(self, batch, get_loss):
        P = _pyngp()
        if self._pairs is None:
            raise RuntimeError("SDF training needs supplied pairs (override_sdf_training_data): online sampling from the mesh is outside the built scope")
        P.check(P.lib().ngpb_model_train_sdf(self._m, batch, int(get_loss)))
