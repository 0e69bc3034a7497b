"""Camera-geometry helpers of the reference API surface, restated on numpy (host side, a handful of values per call).

    pixel_to_camera   monoloco/utils/camera.py:10-29     back-projection  [u, v, 1] K^-T * z_met
    get_keypoints     monoloco/utils/camera.py:69-107    centre / bottom / head / shoulder / hip / ankle of a pose
    xyz_from_distance monoloco/utils/camera.py:161-177   distance along the pixel ray -> xyz

Callers such as `Loco.post_process` use them on one image's detections; the per-detection hot path (pre-process, decode,
bbox-centre rays) runs inside the fused CUDA kernels, the batched pre-process has its own kernel
(`monoloco_b200.engine.preprocess_device`) and whole-split post-processing is `network.post.post_process_batch`.
Inputs may be lists, numpy arrays or torch tensors; results are float32 torch tensors like the reference's."""
import numpy as np
import torch

_F32 = np.float32
# keypoint index ranges of the COCO skeleton parts the reference averages (camera.py:95-105)
_PARTS = {'head': (0, 5), 'shoulder': (5, 7), 'hip': (11, 13), 'ankle': (15, 17)}


def _f32(a):
    if torch.is_tensor(a):
        a = a.detach().cpu().numpy()
    return np.asarray(a, dtype=_F32)


def pixel_to_camera(uv_tensor, kk, z_met):
    """Pixel coordinates (m,2) / (m,x,2) / (m,2,x) -> camera coordinates scaled to depth `z_met`."""
    uv = _f32(uv_tensor)
    if uv.shape[-1] != 2:                      # (m, 2, x): bring the coordinate pair last
        uv = np.swapaxes(uv, 1, 2)
        assert uv.shape[-1] == 2, "Tensor size not recognized"
    k_inv = np.linalg.inv(_f32(kk)).astype(_F32)
    homog = np.concatenate([uv, np.ones(uv.shape[:-1] + (1,), dtype=_F32)], axis=-1)
    return torch.from_numpy(np.ascontiguousarray((homog @ k_inv.T).astype(_F32) * _F32(z_met)))


def get_keypoints(keypoints, mode):
    """One representative pixel per pose: (m,3,17) or (3,17) keypoints -> (m,2)."""
    kps = _f32(keypoints)
    if kps.ndim == 2:
        kps = kps[None]
    assert kps.ndim == 3 and kps.shape[1] == 3, "tensor dimensions not recognized"
    assert mode in ['center', 'bottom', 'head', 'shoulder', 'hip', 'ankle']
    uv = kps[:, 0:2, :]
    if mode in _PARTS:
        lo, hi = _PARTS[mode]
        out = uv[:, :, lo:hi].sum(axis=2, dtype=_F32) / _F32(hi - lo)
    else:
        top, low = uv.max(axis=2), uv.min(axis=2)
        out = (top - low) / _F32(2) + low       # middle of the keypoints' bounding box
        if mode == 'bottom':                    # bottom centre for the KITTI evaluation
            out = np.stack([out[:, 0], top[:, 1]], axis=1)
    return torch.from_numpy(np.ascontiguousarray(out.astype(_F32)))


def xyz_from_distance(distances, xy_centers):
    """xyz at `distances` (float | (m,) | (m,1)) along the rays through normalised image points (m,3) or (3,)."""
    d = _f32(distances).reshape(-1, 1)
    rays = _f32(xy_centers)
    if rays.ndim == 1:
        rays = rays[None]
    assert rays.shape[-1] == 3 and d.shape[-1] == 1, "Size of tensor not recognized"
    norm = np.sqrt(_F32(1) + rays[:, 0:1] ** 2 + rays[:, 1:2] ** 2, dtype=_F32)
    return torch.from_numpy(np.ascontiguousarray((rays * d / norm).astype(_F32)))
