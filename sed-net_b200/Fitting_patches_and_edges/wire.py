"""The text files the stage-1 driver writes for stage 2 (reference generate_predictions_aug.py:424-437, read back by
Fitting_patches_and_edges/primitive_forward_v2.py:1096-1176): per shape id
    {id}_GT_points.txt  N rows 'x;y;z;nx;ny;nz'  '%0.4f'
    {id}_inst.txt       N rows, instance label   '%d'
    {id}_type.txt       N rows, primitive type   '%d'
    {id}_edge.txt       N rows 'p0;p1' softmax of the edge logits '%0.4f'   (optional)
Host-side I/O only (numpy.savetxt / loadtxt, as the reference)."""
import os

import numpy as np


def write_stage1(directory, shape_id, points, normals, inst, types, edge_logits=None):
    """points, normals (N,3); inst, types (N,); edge_logits (2,N) or None.  Returns the paths written."""
    os.makedirs(directory, exist_ok=True)
    base = os.path.join(directory, str(shape_id))
    paths = [base + "_GT_points.txt", base + "_inst.txt", base + "_type.txt"]
    np.savetxt(paths[0], np.concatenate((np.asarray(points), np.asarray(normals)), axis=-1), fmt="%0.4f", delimiter=";")
    np.savetxt(paths[1], np.asarray(inst), fmt="%d")
    np.savetxt(paths[2], np.asarray(types), fmt="%d")
    if edge_logits is not None:
        e = np.asarray(edge_logits, np.float64)
        e = np.exp(e - e.max(0, keepdims=True))
        e = (e / e.sum(0, keepdims=True)).T                     # torch.softmax(edges, dim=1).transpose(1, 2) (:432)
        paths.append(base + "_edge.txt")
        np.savetxt(paths[-1], e.astype(np.float32), fmt="%0.4f", delimiter=";")
    return paths


def read_stage1(directory, shape_id):
    """-> dict(points (N,3) f32, normals (N,3) f32, inst (N,) i64, types (N,) i64[, edges (N,2) f32])."""
    base = os.path.join(directory, str(shape_id))
    pn = np.loadtxt(base + "_GT_points.txt", delimiter=";").astype(np.float32)
    out = dict(points=pn[:, :3], normals=pn[:, 3:6], inst=np.loadtxt(base + "_inst.txt").astype(np.int64),
               types=np.loadtxt(base + "_type.txt").astype(np.int64))
    if os.path.exists(base + "_edge.txt"):
        out["edges"] = np.loadtxt(base + "_edge.txt", delimiter=";").astype(np.float32)
    return out
