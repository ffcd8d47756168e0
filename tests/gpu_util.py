"""Helpers for the -m gpu tests: torch is plumbing only (device memory), every compute call goes through the C ABI."""
import ctypes as C

import numpy as np
import torch

import pyngp


_keepalive = []


def dev(a):
    """numpy -> cuda tensor holding the same bytes (fp16 as int16 views stay fp16). The tensor is kept alive until the
    end of the test (release()), so `ptr(dev(x))` is safe: a dropped temporary would hand its block back to torch's
    caching allocator, and the next dev() in the same argument list would overwrite it before the kernel runs."""
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    _keepalive.append(t)
    return t


def release():
    torch.cuda.synchronize()
    _keepalive.clear()


def ptr(t):
    return C.c_void_p(t.data_ptr())


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


def images_to_device(scene, lens=None):
    """Uploads the scene; returns (device Image array tensor, n, keepalive)."""
    imgs = scene["images"]
    n = len(imgs)
    pix = dev(np.ascontiguousarray(imgs))
    arr = (pyngp.Image * n)()
    per = imgs[0].shape[0] * imgs[0].shape[1] * 4
    for i in range(n):
        arr[i].pixels = pix.data_ptr() + i * per
        arr[i].h, arr[i].w = imgs[i].shape[0], imgs[i].shape[1]
        arr[i].fx, arr[i].fy, arr[i].cx, arr[i].cy = scene["fx"], scene["fy"], scene["cx"], scene["cy"]
        if lens is not None:  # (ELensMode, 7 parameters, principal point) for every image
            arr[i].lens_mode = int(lens[0])
            for k in range(7):
                arr[i].lens_params[k] = float(lens[1][k])
            arr[i].cx, arr[i].cy = lens[2]
        cm = np.asarray(scene["xforms"][i], dtype=np.float32).reshape(3, 4).T.reshape(-1).copy()
        eff = np.empty(12, np.float32)
        pyngp.lib().ngpb_effective_xform(cm.ctypes.data_as(C.c_void_p), eff.ctypes.data_as(C.c_void_p))
        for k in range(12):
            arr[i].raw_xform[k] = float(cm[k])
            arr[i].xform[k] = float(eff[k])
    raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
    meta = dev(raw)
    return meta, n, (pix, arr)


def rng_struct(orc_rng):
    return pyngp.Rng(orc_rng.state, orc_rng.inc)
