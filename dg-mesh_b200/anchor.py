"""Gaussian <-> mesh anchoring on the device (the third member of densify / prune / anchor).

Drop-in for `GaussianModelDPSRDynamicAnchor.anchor_mesh`
(dgmesh/scene/gaussian_model_dpsr_dynamic_anchor.py:736-828; called from train.py:293 every anchor_interval
iterations after anchor_iter).  Same arguments, same return value (the anchor loss, differentiable w.r.t. the
deformation network exactly as the reference's), same random-number consumption (two `torch.randperm`, then
`densify_from_face`'s `randn`), same resulting model:

  reference                                                      here
  -------------------------------------------------------------  ------------------------------------------------
  trimesh on the host for centroids / normals (D2H + H2D)        a few torch ops in float64 on the device
  pytorch3d.ops.knn_points(K=1)                                  dgk_nearest (csrc/knn.cu), distance re-formed
                                                                 differentiably from the returned index
  torch.unique + three torch.isin over all faces                 one bincount
  [bs, G, 1] match mask, cumsum over G (256 x G int64 = 0.4 GB   one stable sort of the selected faces' members;
  at 200k), xor, masked_select over [bs, G]                      ranks from a bincount prefix
  average_and_prune on the [bs, G, 1] mask                       the same averaging on a [bs, topn] index table

`prune_points`, `densification_postfix` and `densify_from_face` are the model's own methods (the optimiser
surgery is shared with densify_and_prune; dg-mesh_b200/densify.py replaces the hot one)."""
import os
import sys

import torch
import torch.nn as nn

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
import _dgm_lib  # noqa: E402


def nearest(queries, refs):
    """(squared distance [Q], index [Q] int64) of the nearest reference point of every query (no gradient)."""
    if not (queries.is_cuda and refs.is_cuda):
        raise ValueError("anchor.nearest: CUDA tensors required (no CPU fallback)")
    q = queries.detach().contiguous().float()
    r = refs.detach().contiguous().float()
    if q.dim() != 2 or q.shape[1] != 3 or r.dim() != 2 or r.shape[1] != 3 or r.shape[0] < 1:
        raise ValueError("anchor.nearest: expected queries [Q,3] and refs [R>=1,3]")
    d2 = torch.empty((q.shape[0],), dtype=torch.float32, device=q.device)
    idx = torch.empty((q.shape[0],), dtype=torch.int64, device=q.device)
    _dgm_lib.check(_dgm_lib.lib().dgk_nearest(q.shape[0], q.data_ptr(), r.shape[0], r.data_ptr(), d2.data_ptr(),
                                              idx.data_ptr(), _dgm_lib.stream_ptr()), "dgk_nearest")
    return d2, idx


def face_geometry(verts, faces):
    """Centroids, unit normals [F,3] fp32 and the mean unique-edge length, computed in float64 like trimesh
    (`triangles_center`, `face_normals`, `edges_unique_length.mean()`)."""
    v = verts.detach().double()
    f = faces.long()
    tri = v[f]                                                        # [F,3,3]
    centroids = tri.mean(1)
    cross = torch.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0], dim=-1)
    normals = cross / cross.norm(dim=-1, keepdim=True).clamp_min(1e-300)
    e = torch.cat([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0)
    e = torch.unique(torch.sort(e, dim=1).values, dim=0)
    edge_len = (v[e[:, 0]] - v[e[:, 1]]).norm(dim=-1).mean()
    return centroids.float(), normals.float(), edge_len


def select_topn(face_indices, faces, n_faces, topn):
    """Which Gaussians of the drawn faces survive to be averaged, and which are deleted.

    face_indices [G] int64: the face every Gaussian is assigned to; faces [X]: the drawn faces (distinct).
    Returns (to_delete [G] bool, members [X, topn] int64): per drawn face, in the order of `faces`, its first `topn`
    Gaussians by index; every further Gaussian of a drawn face is marked for deletion.  Same selection as the
    reference's `[X, G, 1]` match mask + row-wise cumsum (gaussian_model_dpsr_dynamic_anchor.py:787-799), obtained
    from one stable sort of the drawn faces' Gaussians.  Raises if a drawn face has fewer than `topn` Gaussians
    (the reference's `masked_select(...).view(-1, topn, .)` fails there too)."""
    dev = face_indices.device
    X = faces.shape[0]
    row_of_face = torch.full((n_faces,), -1, dtype=torch.int64, device=dev)
    row_of_face[faces] = torch.arange(X, device=dev)
    row = row_of_face[face_indices]                                     # [G] row in the batch, -1: not drawn
    member = torch.nonzero(row >= 0).squeeze(1)                         # ascending Gaussian index
    order = torch.sort(row[member], stable=True).indices               # by row, index order kept inside a row
    member = member[order]
    rows_sorted = row[member]
    per_row = torch.bincount(rows_sorted, minlength=X)
    first = torch.cumsum(per_row, 0) - per_row
    rank = torch.arange(member.shape[0], device=dev) - first[rows_sorted]
    keep = rank < topn
    if X and int(per_row.min()) < topn:
        raise RuntimeError("anchor_mesh: a drawn face has fewer than topn Gaussians")
    to_delete = torch.zeros(face_indices.shape[0], dtype=torch.bool, device=dev)
    to_delete[member[~keep]] = True
    return to_delete, member[keep].view(X, topn)


def _average_and_prune(self, sel, deform, deform_back, t):
    """`average_and_prune` (:599-649) on an index table sel [X, topn] instead of a [X, G, 1] mask."""
    X, topn = sel.shape
    idx = sel.reshape(-1)
    pick = lambda p: p[idx].view(X, topn, *p.shape[1:])     # noqa: E731
    selected_xyz, selected_scaling = pick(self._xyz), pick(self._scaling)
    selected_rotation, selected_normal = pick(self._rotation), pick(self._normal)
    new_features_dc = pick(self._features_dc).mean(1, keepdim=True)
    new_features_rest = pick(self._features_rest).mean(1, keepdim=True)
    new_opacity = pick(self._opacity).mean(1, keepdim=True)
    with torch.no_grad():
        time_input = torch.ones(X * topn, 1, device="cuda") * t
        d_xyz, d_rotation, d_scaling, d_normal = deform.step(selected_xyz.view(-1, 3), time_input)
        selected_xyz = selected_xyz + d_xyz.view(X, topn, -1)
        selected_scaling = selected_scaling + d_scaling.view(X, topn, -1)
        selected_rotation = selected_rotation + d_rotation.view(X, topn, -1)
        selected_normal = selected_normal + d_normal.view(X, topn, -1)
    new_xyz = selected_xyz.mean(1, keepdim=True)
    deformed_xyz = new_xyz
    new_scaling = selected_scaling.mean(1, keepdim=True)
    new_rotation = selected_rotation.mean(1, keepdim=True)
    new_normal = selected_normal.mean(1, keepdim=True)
    with torch.no_grad():
        time_input = torch.ones(X, 1, device="cuda") * t
        d_xyz, d_rotation, d_scaling, d_normal = deform_back.step(new_xyz.view(-1, 3), time_input)
        new_xyz = new_xyz.view(X, -1) + d_xyz
        new_scaling = new_scaling.view(X, -1) + d_scaling
        new_rotation = new_rotation.view(X, -1) + d_rotation
        new_normal = nn.functional.normalize(new_normal.view(X, -1) + d_normal, p=2, dim=-1)
    gone = torch.zeros(self._xyz.shape[0], dtype=torch.bool, device=idx.device)
    gone[idx] = True
    self.prune_points(gone)
    self.densification_postfix(new_xyz, new_features_dc.squeeze(1), new_features_rest.squeeze(1),
                               new_opacity.squeeze(1), new_scaling, new_rotation, new_normal)
    return deformed_xyz.view(-1, 3)


def anchor_mesh(self, verts, faces, deform, deform_back, t, search_radius=0.0005, topn=2, bs=256, increase_bs=1024):
    search_radius = self.gaussian_scale * search_radius
    old_xyz_num = self.get_xyz.shape[0]
    dev = self.get_xyz.device

    time_input = torch.ones(old_xyz_num, 1, device="cuda") * t
    d_xyz, _, _, _ = deform.step(self.get_xyz.detach(), time_input)
    centroids, normals, avg_edge_length = face_geometry(verts, faces)
    n_faces = centroids.shape[0]
    gaussian_points = self.get_xyz + d_xyz

    # every Gaussian -> its closest face centroid; the squared distance is re-formed from the index so that it
    # carries the gradient pytorch3d's knn_points gives it (d/dp |p - c|^2 at the fixed neighbour)
    d2, face_indices = nearest(gaussian_points, centroids)
    gs_face_dist = ((gaussian_points - centroids[face_indices]) ** 2).sum(-1, keepdim=True)
    valid = d2 < search_radius
    self.prune_points(~valid)
    invalid_ratio = (~valid).sum().item() / face_indices.shape[0]
    gs_face_dist = gs_face_dist[valid]
    face_indices = face_indices[valid]

    # faces with exactly one / several / no Gaussian
    counts = torch.bincount(face_indices, minlength=n_faces)
    face_indices_n_1 = torch.nonzero(counts > 1).squeeze(1)
    n_1_1 = int((counts == 1).sum())
    anchor_loss_1_1 = gs_face_dist[counts[face_indices] == 1].mean()

    # n-1: a random batch of such faces; per face the first `topn` Gaussians (by index) are averaged into one new
    # Gaussian, the others are deleted
    random_indices = torch.randperm(face_indices_n_1.shape[0], device=dev)[:bs]
    face_indices_n_1 = face_indices_n_1[random_indices]
    to_delete, members = select_topn(face_indices, face_indices_n_1, n_faces, topn)
    self.prune_points(to_delete)
    new_index = torch.cumsum(~to_delete, 0) - 1                          # positions after the prune
    sel = new_index[members]
    new_xyz = _average_and_prune(self, sel, deform, deform_back, t)
    face_xyz = centroids[face_indices_n_1]
    anchor_loss_n_1 = torch.norm(face_xyz - new_xyz, dim=-1).mean()

    # 0-1: faces neither matched one-to-one nor in the batch above get a fresh Gaussian at their centroid
    # (as in the reference this includes the n-1 faces that were not drawn)
    selected = torch.zeros(n_faces, dtype=torch.bool, device=dev)
    selected[face_indices_n_1] = True
    face_0_1 = (counts != 1) & ~selected
    c0, n0 = centroids[face_0_1], normals[face_0_1]
    random_indices = torch.randperm(c0.shape[0], device=dev)[:increase_bs]
    self.densify_from_face(c0[random_indices], n0[random_indices], avg_edge_length / 2, deform_back, t)

    anchor_loss = anchor_loss_1_1 + anchor_loss_n_1
    new_xyz_num = self.get_xyz.shape[0]
    print(f"Old number of gaussians: {old_xyz_num}, New number of gaussians: {new_xyz_num}, Target face number "
          f"{n_faces}, Anchor loss: {anchor_loss:04f} 1-1 Hit rate: {n_1_1 / n_faces:.4f} "
          f"Invalid ratio: {invalid_ratio:.4f}")
    return anchor_loss


def install(model_class):
    """Replace `anchor_mesh` on a reference Gaussian-model class; returns the original method."""
    orig = model_class.anchor_mesh
    if orig is anchor_mesh:
        return getattr(model_class, "_reference_anchor_mesh", None)
    model_class._reference_anchor_mesh = orig
    model_class.anchor_mesh = anchor_mesh
    return orig
