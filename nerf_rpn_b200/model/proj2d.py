"""2-D projection regression loss of the reference (rpn.py:28-106, 421-453 `loss_rpn_box_reg_2d`, weight `--reg_loss_weight_2d`;
fcos/loss.py:452-485 `compute_2d_projection_loss`, weight `--proj2d_loss_weight`; fcos/utils.py:300-377): two characteristic points of every
predicted / target box (an AABB's min and max corner; an OBB's centre -+ its rotated half diagonal, `obb2points_3d`) are projected into four
pinhole cameras that look at the scene's centroid from its upper corners, and the projections are compared with a smooth-L1 (beta 1/9).

Training only, on the sampled positives of a step (<= 128 anchors per mesh for the RPN, the positive locations for FCOS): a few dozen
element-wise torch ops whose autograd gradient the training engine adds to d(pred) -- the same arrangement as the IoU-type losses
(DESIGN.md section 1).  Both weights are 0 in every recipe the reference ships (train.sh, train_fcos.sh)."""
import numpy as np
import torch

IMG_W, IMG_H, FOCAL_X, FOCAL_Y = 640, 480, 600, 600          # rpn.py:422, fcos/loss.py:454
BETA = 1.0 / 9


def w2c_matrices(res) -> np.ndarray:
    """(4, 4, 4) fp32 world-to-camera matrices (get_w2cs): cameras at centroid + (+-res, +-res, res), z axis pointing from the centroid to the
    camera, world up = +z; built in fp64, inverted, rounded once like torch.Tensor(np.linalg.inv(...))."""
    centroid = np.full(3, res / 2.0)
    out = []
    for sx, sy in ((1, 1), (1, -1), (-1, 1), (-1, -1)):
        cam = centroid + np.array([sx * res, sy * res, res], np.float64)
        z = cam - centroid
        z = z / np.linalg.norm(z)
        x = np.cross(np.array([0.0, 0.0, 1.0]), z)
        x = x / np.linalg.norm(x)
        y = np.cross(z, x)
        y = y / np.linalg.norm(y)
        c2w = np.eye(4)
        c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = x, y, z, cam
        out.append(np.linalg.inv(c2w))
    return np.stack(out).astype(np.float32)


_cache = {}


def cameras(res, device):
    """-> (poses (4,4,4), intrinsics (3,3)) on `device`, cached per (res, device)."""
    key = (float(res), str(device))
    if key not in _cache:
        k = torch.tensor([[FOCAL_X, 0, IMG_W / 2], [0, FOCAL_Y, IMG_H / 2], [0, 0, 1]], dtype=torch.float32, device=device)
        _cache[key] = (torch.from_numpy(w2c_matrices(res)).to(device), k)
    return _cache[key]


def box_points(boxes: torch.Tensor) -> torch.Tensor:
    """(K, 6) [x1,y1,z1,x2,y2,z2] -> (2K, 3): the K min corners, then the K max corners (rpn.py:430-433);
    (K, 7) [x,y,z,w,l,h,theta] -> (2K, 3): centre - v, then centre + v, v = (w/2 cos - l/2 sin, w/2 sin + l/2 cos, h/2) (obb2points_3d)."""
    if boxes.shape[1] == 6:
        return torch.cat([boxes[:, :3], boxes[:, 3:]], 0)
    c, w, l, h, th = boxes[:, :3], boxes[:, 3], boxes[:, 4], boxes[:, 5], boxes[:, 6]
    co, si = torch.cos(th), torch.sin(th)
    v = torch.stack([w / 2 * co - l / 2 * si, w / 2 * si + l / 2 * co, h / 2], 1)
    return torch.cat([c - v, c + v], 0)


def project_all(points: torch.Tensor, res) -> torch.Tensor:
    """(M, 3) world points -> (4M, 2) pixel coordinates, camera by camera (project + the loop over pose_list)."""
    poses, k = cameras(res, points.device)
    hom = torch.cat([points, torch.ones_like(points[:, :1])], 1)
    out = []
    for pose in poses:
        cam = hom @ pose.t()
        pic = cam[:, :3] @ k.t()
        out.append(pic[:, :2] / pic[:, 2:3])
    return torch.cat(out, 0)


def smooth_l1(a: torch.Tensor, b: torch.Tensor, beta: float = BETA) -> torch.Tensor:
    d = (a - b).abs()
    return torch.where(d < beta, 0.5 * d * d / beta, d - 0.5 * beta)


def rpn_projection_loss(pred_boxes: torch.Tensor, target_boxes: torch.Tensor, n_positive: int, max_mesh_dim) -> torch.Tensor:
    """rpn.py:421-453: sum smooth-L1 of the projected points / sampled positives / max mesh dimension.  Differentiable w.r.t. pred_boxes."""
    p2 = project_all(box_points(pred_boxes), max_mesh_dim)
    t2 = project_all(box_points(target_boxes), max_mesh_dim)
    return smooth_l1(p2, t2).sum() / n_positive / max_mesh_dim


def fcos_projection_loss(pred_boxes: torch.Tensor, target_boxes: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """fcos/loss.py:452-485 (cameras of a 160-voxel scene): the centerness-weighted mean over the 8 projected points x 2 coordinates of a box,
    summed over the boxes -- before the division by sum_centerness_targets_avg_per_gpu."""
    p2 = project_all(box_points(pred_boxes), 160)
    t2 = project_all(box_points(target_boxes), 160)
    per = smooth_l1(p2, t2) / 160
    factor = per.shape[0] // weights.shape[0]
    return (per * weights.repeat(factor)[:, None]).sum() / (factor * per.shape[1])
