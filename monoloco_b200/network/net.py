"""
`Loco` -- the drop-in for monoloco/network/net.py:23-133 (same constructor, `forward`, `post_process`,
attributes), running pre-process + network + decode as one fused CUDA kernel per image.

Differences that are deliberate and documented:
  * the result tensors are produced on the GPU and returned as CPU float32 tensors, exactly the dict layout
    callers index (`dic_out['xyzd'][:, 0:3]`, `dic_out['yaw'][0][idx]`, ...), plus one extra key `xyz_c` =
    xyz_from_distance(d, bbox-centre ray) that `post_process` reuses;
  * MC-dropout epistemic std uses an in-kernel counter RNG + inverse-CDF Laplace sampling instead of torch's
    reseeded generator (equal in distribution, not sample-for-sample);
  * there is no CPU mode: `device=None` selects the current CUDA device.
"""
from collections import defaultdict

import torch

from .. import _lib as L_
from ..engine import dec_to_dict
from ..utils import get_iou_matches, reorder_matches, get_keypoints, pixel_to_camera, xyz_from_distance
from ..activity import social_interactions, is_raising_hand
from .architectures import MonolocoModel, LocoModel


class Loco:
    """Class for both MonoLoco and MonStereo (net.py:23-81)."""
    LINEAR_SIZE_MONO = 256
    N_SAMPLES = 100

    def __init__(self, model, mode, net=None, device=None, n_dropout=0, p_dropout=0.2, linear_size=1024):
        assert mode in ('mono', 'stereo'), "mode not recognized"
        self.mode = mode
        if net is None:
            self.net = 'monoloco_pp' if mode == 'mono' else 'monstereo'
        else:
            # net.py:41-44 is unreachable in the reference (reads self.net before assignment); here the documented
            # intent is implemented: legacy nets are mono-only
            assert net in ('monstereo', 'monoloco', 'monoloco_p', 'monoloco_pp')
            if net != 'monstereo':
                assert mode == 'mono', "Assert arguments mode and net are in conflict"
            self.net = net

        if self.net == 'monstereo':
            input_size, output_size = 68, 10
        elif self.net == 'monoloco_p':
            input_size, output_size, linear_size = 34, 9, 256
        elif self.net == 'monoloco_pp':
            input_size, output_size = 34, 9
        else:
            input_size, output_size = 34, 2

        if not torch.cuda.is_available():
            raise RuntimeError("monoloco_b200.Loco needs a CUDA device (B200); there is no CPU fallback")
        self.device = torch.device('cuda', torch.cuda.current_device()) if not device else torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError("monoloco_b200.Loco runs on CUDA devices only")
        self.n_dropout = n_dropout
        self.epistemic = bool(self.n_dropout > 0)

        if isinstance(model, str):
            if self.net in ('monoloco', 'monoloco_p'):
                self.model = MonolocoModel(p_dropout=p_dropout, input_size=input_size, linear_size=linear_size,
                                           output_size=output_size)
            else:
                self.model = LocoModel(p_dropout=p_dropout, input_size=input_size, output_size=output_size,
                                       linear_size=linear_size, device=self.device)
            self.model.load_state_dict(torch.load(model, map_location=lambda storage, loc: storage))
        else:
            self.model = model
        self.model.eval()
        self.model.to(self.device)

    # ------------------------------------------------------------------------------------------- forward
    def forward(self, keypoints, kk, keypoints_r=None):
        """net.py:83-133: keypoints [m][3][17] lists (+ right keypoints for stereo) -> dict of CPU tensors."""
        if not keypoints:
            return None
        eng = self.model.engine()
        with torch.no_grad():
            if self.net == 'monstereo':
                kps = torch.tensor(keypoints, dtype=torch.float32).to(self.device)
                if keypoints_r:
                    kps_r = torch.tensor(keypoints_r, dtype=torch.float32).to(self.device)
                else:
                    kps_r = kps[0:1, :].clone()  # net.py:115-116
                out = eng.forward(kps, x_right=kps_r, kk=kk, kind=L_.IN_KPS_STEREO, want_xyzc=True)
                raw, dec, xyzc = eng.stereo_filter_host(out['raw'], out['dec'], out['xyzc'], kps.shape[0],
                                                        kps_r.shape[0])  # process.py:307-327, one sync
                dic_out = dec_to_dict(raw, dec, stereo=True)
                dic_out['xyz_c'] = xyzc[:, 0:3]
                n_out = kps.shape[0]  # net.py:130: outputs is the clustered 3-D tensor -> number of left poses
                inputs = None
            elif not self.epistemic and self.net != 'monoloco':
                # per-image fast path: one C call does H2D (pinned staging), the fused kernel, D2H and the sync
                out = self._forward_host_mono(eng, keypoints, kk)
                raw, dec = out['raw'], out['dec']
                if self.net == 'monoloco_p':
                    r, d = raw, dec
                    dic_out = {'xyz': r[:, 0:3], 'zb': r[:, 2:4], 'h': r[:, 4:5], 'w': r[:, 5:6], 'l': r[:, 6:7],
                               'ori': r[:, 7:9], 'xyzd': d[:, 0:4], 'd': d[:, 3:4], 'bi': d[:, 4:5],
                               'yaw': (d[:, 5:6], d[:, 6:7])}
                else:
                    dic_out = dec_to_dict(raw, dec, stereo=False)
                dic_out['xyz_c'] = out['xyzc'][:, 0:3]
                dic_out['epi'] = [0.] * raw.shape[0]
                return dic_out
            else:
                kps = torch.tensor(keypoints, dtype=torch.float32).to(self.device)
                zero_center = self.net == 'monoloco'
                out = eng.forward(kps, kk=kk, kind=L_.IN_KPS, want_xyzc=True, want_x=self.epistemic,
                                  zero_center=zero_center)
                raw, dec = out['raw'], out['dec']
                if self.net == 'monoloco':
                    dic_out = {'d': raw[:, 0:1].cpu(), 'bi': dec[:, 4:5].cpu()}  # net.py:95-100
                elif self.net == 'monoloco_p':
                    r, d = raw.cpu(), dec.cpu()  # extract_outputs_mono, process.py:330-360
                    dic_out = {'xyz': r[:, 0:3], 'zb': r[:, 2:4], 'h': r[:, 4:5], 'w': r[:, 5:6], 'l': r[:, 6:7],
                               'ori': r[:, 7:9], 'xyzd': d[:, 0:4], 'd': d[:, 3:4], 'bi': d[:, 4:5],
                               'yaw': (d[:, 5:6], d[:, 6:7])}
                else:
                    dic_out = dec_to_dict(raw, dec, stereo=False)
                dic_out['xyz_c'] = out['xyzc'][:, 0:3].cpu()
                n_out = raw.shape[0]
                inputs = out.get('x')
            if self.n_dropout > 0 and self.net != 'monstereo':
                dic_out['epi'] = self.epistemic_uncertainty(inputs)
            else:
                dic_out['epi'] = [0.] * n_out
        return dic_out

    def _forward_host_mono(self, eng, keypoints, kk):
        """Python lists -> pinned staging -> mlb_forward_host -> fresh CPU tensors (caller owns them, net.py contract)."""
        import numpy as np
        m = len(keypoints)
        st = getattr(self, '_stage', None)
        if st is None or st['cap'] < m:
            cap = max(64, 1 << (m - 1).bit_length())
            st = {'cap': cap, 'kps': torch.empty((cap, 3, 17), dtype=torch.float32).pin_memory(),
                  'raw': torch.empty((cap, eng.output_size), dtype=torch.float32).pin_memory(),
                  'dec': torch.empty((cap, 8), dtype=torch.float32).pin_memory(),
                  'xyzc': torch.empty((cap, 4), dtype=torch.float32).pin_memory()}
            self._stage = st
        st['kps'][:m] = torch.from_numpy(np.asarray(keypoints, dtype=np.float32))
        out = {'raw': st['raw'][:m], 'dec': st['dec'][:m], 'xyzc': st['xyzc'][:m]}
        eng.forward_host(st['kps'][:m], kk=kk, kind=L_.IN_KPS, out=out)
        return {k: v.clone() for k, v in out.items()}

    def epistemic_uncertainty(self, inputs):
        """net.py:135-161: n_dropout stochastic passes (top-level dropout on) + Laplace sampling -> std per instance."""
        assert self.net in ('monoloco', 'monoloco_p', 'monoloco_pp'), "Not supported for MonStereo"
        eng = self.model.engine()
        return eng.epistemic_std(inputs, self.n_dropout, n_samples=self.N_SAMPLES, seed=1).cpu()

    # ------------------------------------------------------------------------------------------- post-process
    @staticmethod
    def post_process(dic_in, boxes, keypoints, kk, dic_gt=None, iou_min=0.3, reorder=True, verbose=False):
        """Final per-instance dictionary for visualisation / KITTI txt (same keys, order and values as net.py:163-248),
        computed column-wise: instances matched to ground truth come first (left to right), then the rest."""
        import numpy as np
        res = defaultdict(list)
        if dic_in is None:
            return res
        n = len(boxes)
        matches = get_iou_matches(boxes, dic_gt['boxes'], iou_min=iou_min) if dic_gt else []
        if verbose:
            print("found {} matches with ground-truth".format(len(matches)) if dic_gt else "NO ground-truth associated")
        taken = {i for i, _ in matches}
        if reorder and matches:
            matches = reorder_matches(matches, boxes, mode='left_right')
        order = [i for i, _ in matches] + [i for i in range(n) if i not in taken]
        res['gt'] = [True] * len(matches) + [False] * (n - len(matches))

        col = lambda key: np.asarray(dic_in[key], dtype=np.float64).reshape(-1)  # noqa: E731
        dd, bi, epi = col('d'), col('bi'), np.asarray(dic_in['epi'], dtype=np.float64).reshape(-1)
        centres = get_keypoints(keypoints, mode='center')
        rays = pixel_to_camera(centres, kk, 1)  # bbox-centre rays at z = 1 (net.py:195)
        if isinstance(dic_in, dict) and dic_in.get('xyz_c') is not None and len(dic_in['xyz_c']) == n:
            xyz = dic_in['xyz_c']  # xyz_from_distance already evaluated by the fused kernel's epilogue
        else:
            xyz = xyz_from_distance(torch.as_tensor(dd, dtype=torch.float32), rays)
        xyz64 = np.asarray(xyz, dtype=np.float64)
        conf = 0.035 * np.asarray([b[-1] for b in boxes], dtype=np.float64) / (bi / np.sqrt((xyz64 ** 2).sum(1)))
        pix = {name: np.asarray(get_keypoints(keypoints, mode=mode)).tolist()
               for name, mode in (('uv_centers', 'center'), ('uv_shoulders', 'shoulder'), ('uv_heads', 'head'))}
        has_yaw, has_aux = 'yaw' in dic_in, 'aux' in dic_in
        for i in order:
            res['boxes'].append(boxes[i])
            res['confs'].append(float(conf[i]))
            res['dds_pred'].append(float(dd[i]))
            res['stds_ale'].append(float(bi[i]))
            res['stds_epi'].append(float(epi[i]))
            res['xyz_pred'].append(np.asarray(xyz[i]).reshape(-1).tolist())
            res['uv_kps'].append(keypoints[i])
            for name, pts in pix.items():
                res[name].append([round(pts[i][0]), round(pts[i][1])])
            res['angles']  # the reference's defaultdict access creates these keys even when the value is missing
            if not has_yaw:
                continue
            res['angles'].append(float(dic_in['yaw'][0][i]))
            res['angles_egocentric'].append(float(dic_in['yaw'][1][i]))
            res['aux']
            if has_aux:
                res['aux'].append(float(dic_in['aux'][i]))
        for i, j in matches:
            d_real = dic_gt['ys'][j][3]
            res['dds_real'].append(d_real)
            res['boxes_gt'].append(dic_gt['boxes'][j])
            res['xyz_real'].append(xyz_from_distance(d_real, rays[i]).squeeze().tolist())
        return res

    @staticmethod
    def social_distance(dic_out, args):
        """net.py:250-265: flag every instance whose F-formation test fires."""
        angles, dds, stds = dic_out['angles'], dic_out['dds_pred'], dic_out['stds_ale']
        xz_centers = [[xx[0], xx[2]] for xx in dic_out['xyz_pred']]
        dic_out['social_distance'] = [bool(social_interactions(idx, xz_centers, angles, dds, stds=stds,
                                                               threshold_prob=args.threshold_prob,
                                                               threshold_dist=args.threshold_dist, radii=args.radii))
                                      for idx, _ in enumerate(dic_out['xyz_pred'])]
        return dic_out

    @staticmethod
    def raising_hand(dic_out, keypoints):
        """net.py:267-271."""
        dic_out['raising_hand'] = [is_raising_hand(keypoint) for keypoint in keypoints]
        return dic_out
