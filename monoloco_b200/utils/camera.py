"""Small camera-geometry helpers of the reference API surface (monoloco/utils/camera.py:10-29, 69-107, 161-177).

These are the *stand-alone* helpers callers such as Loco.post_process use on a handful of values; the
per-detection hot path (pre-process / decode) runs inside the fused CUDA kernel, and the batched
pre-process has its own kernel (`monoloco_b200.engine.preprocess_device`)."""
import numpy as np
import torch
import torch.nn.functional as F


def pixel_to_camera(uv_tensor, kk, z_met):
    """utils/camera.py:10-29: [u, v, 1] K^-T * z_met for (m,2) / (m,x,2) / (m,2,x) inputs."""
    if isinstance(uv_tensor, (list, np.ndarray)):
        uv_tensor = torch.tensor(uv_tensor)
    if isinstance(kk, (list, np.ndarray)):
        kk = torch.tensor(kk)
    if uv_tensor.size()[-1] != 2:
        uv_tensor = uv_tensor.permute(0, 2, 1)
        assert uv_tensor.size()[-1] == 2, "Tensor size not recognized"
    uv_padded = F.pad(uv_tensor, pad=(0, 1), mode="constant", value=1)
    return torch.matmul(uv_padded, torch.inverse(kk).t()) * z_met


def get_keypoints(keypoints, mode):
    """utils/camera.py:69-107: centre / bottom / head / shoulder / hip / ankle point of (m,3,17) keypoints."""
    if isinstance(keypoints, (list, np.ndarray)):
        keypoints = torch.tensor(keypoints)
    if len(keypoints.size()) == 2:
        keypoints = keypoints.unsqueeze(0)
    assert len(keypoints.size()) == 3 and keypoints.size()[1] == 3, "tensor dimensions not recognized"
    assert mode in ['center', 'bottom', 'head', 'shoulder', 'hip', 'ankle']
    kps_in = keypoints[:, 0:2, :]
    if mode in ('center', 'bottom'):
        kmax, _ = kps_in.max(2)
        kmin, _ = kps_in.min(2)
        if mode == 'center':
            return (kmax - kmin) / 2 + kmin
        return torch.cat(((kmax[:, 0:1] - kmin[:, 0:1]) / 2 + kmin[:, 0:1], kmax[:, 1:2]), -1)
    sl = {'head': slice(0, 5), 'shoulder': slice(5, 7), 'hip': slice(11, 13), 'ankle': slice(15, 17)}[mode]
    return kps_in[:, :, sl].mean(2)


def xyz_from_distance(distances, xy_centers):
    """utils/camera.py:161-177."""
    if isinstance(distances, float):
        distances = torch.tensor(distances).unsqueeze(0)
    if len(distances.size()) == 1:
        distances = distances.unsqueeze(1)
    if len(xy_centers.size()) == 1:
        xy_centers = xy_centers.unsqueeze(0)
    assert xy_centers.size()[-1] == 3 and distances.size()[-1] == 1, "Size of tensor not recognized"
    return xy_centers * distances / torch.sqrt(1 + xy_centers[:, 0:1].pow(2) + xy_centers[:, 1:2].pow(2))
