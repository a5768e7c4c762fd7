"""Output stage of the sampling scripts, on the device.

Reference: sample/generate_uncond.py:113-122 (same in all five scripts)

    pred_mesh_o3d = get_o3d_mesh_from_tensors(v, t)                 utils/utils.py:79-121
    o3d.io.write_triangle_mesh(mesh_path, pred_mesh_o3d)            open3d 0.18.0 .obj writer
    ms = ml.MeshSet(); ms.set_verbosity(False); ms.load_new_mesh(mesh_path)
    ms.apply_coord_laplacian_smoothing()                            pymeshlab 2023.12 (MeshLab "Laplacian Smooth", defaults)
    ms.meshing_remove_connected_component_by_face_number(mincomponentsize=2500)
    ms.save_current_mesh(mesh_path)

open3d and pymeshlab are third-party and absent here: PARITY UNPINNED.  The filters restate the published MeshLab / vcglib
algorithms (see oracle/meshclean_oracle.py for the literal numpy restatement they are tested against): Laplacian smoothing
with cotangent weights, 3 Jacobi steps, border vertices following the border polyline; connected components over
edge-adjacent faces, components with fewer than `mincomponentsize` faces removed, then unreferenced vertices removed.
The mesh stays on the GPU between marching cubes and the file: smoothing is a scatter-add over the face list, the
components are found by min-label hooking + pointer jumping over the shared-edge pairs (O(log F) rounds).
`MeshSet` / `io.write_triangle_mesh` below mirror the call surface the scripts use, so the output block of a script runs
with `import open3d as o3d` / `import pymeshlab as ml` swapped for `from surfd_b200 import output as o3d, output as ml`.
"""
import os

import torch


class TriangleMesh:
    """what get_o3d_mesh_from_tensors returns: vertices float64 [V,3] (open3d stores doubles), triangles int32 [F,3]"""

    def __init__(self, vertices=None, triangles=None):
        self.vertices = vertices
        self.triangles = triangles


def get_o3d_mesh_from_tensors(vertices, triangles):
    """utils/utils.py:79-121 for the (V,3) / (F,3) case the scripts use; the data stays where it is (no .cpu().numpy())"""
    v = torch.as_tensor(vertices)
    t = torch.as_tensor(triangles)
    if v.dim() != 2 or v.shape[1] not in (3, 6, 9) or t.dim() != 2 or t.shape[1] not in (3, 6):
        raise ValueError("vertices must be (NUM_VERTICES, 3|6|9) and triangles (NUM_TRIANGLES, 3|6)")
    if v.shape[1] != 3 or t.shape[1] != 3:
        raise NotImplementedError("normals / colours are never passed by the Surf-D scripts")
    return TriangleMesh(v[:, :3].detach().clone().to(torch.float64), t[:, :3].detach().clone().to(torch.int32))


def _fmt_g(arr):
    """C++ iostream default formatting (6 significant digits, %g) of a float64 numpy array, vectorised"""
    import numpy as np
    return np.char.mod("%g", arr)


def _native_write(path, header, v64, number_format, mid, f64, footer):
    """the text conversion runs in the C library (csrc/objio.cu: surfd_obj_write); v64 float64 [V,3], f64 int64 [F,3] 0-based, host"""
    import ctypes
    import numpy as np
    from . import _lib
    v64 = np.ascontiguousarray(v64, dtype=np.float64).reshape(-1, 3)
    f64 = np.ascontiguousarray(f64, dtype=np.int64).reshape(-1, 3)
    _lib.check(_lib.load().surfd_obj_write(os.fsencode(path), header.encode(), ctypes.c_void_p(v64.ctypes.data), v64.shape[0], int(number_format),
                                           mid.encode(), ctypes.c_void_p(f64.ctypes.data), f64.shape[0], footer.encode()))


def write_obj_o3d(path, vertices, triangles):
    """open3d 0.18.0 WriteTriangleMeshToOBJ layout (cpp/open3d/io/file_format/FileOBJ.cpp) for a mesh without normals,
    colours or uvs: 4 comment lines, `v x y z` with default ostream precision (%g), `f a b c` 1-based"""
    v = vertices.detach().to(torch.float64).cpu().numpy()
    f = triangles.detach().cpu().numpy()
    name = os.path.splitext(os.path.basename(path))[0]
    header = "# Created by Open3D \n# object name: %s\n# number of points: %d\n# number of triangles: %d\n" % (name, len(v), len(f))
    _native_write(path, header, v, 1, "", f, "")
    return True


def _write_obj_o3d_py(path, vertices, triangles):
    """the same layout formatted by numpy / Python: the specification the native writer is tested against (tests only)"""
    import numpy as np
    v = vertices.detach().to(torch.float64).cpu().numpy()
    f = triangles.detach().cpu().numpy().astype(np.int64) + 1
    name = os.path.splitext(os.path.basename(path))[0]
    vs = _fmt_g(v)
    lines = ["# Created by Open3D ", "# object name: " + name, "# number of points: %d" % len(v), "# number of triangles: %d" % len(f)]
    body_v = np.char.add(np.char.add(np.char.add(np.char.add(np.char.add("v ", vs[:, 0]), " "), vs[:, 1]), " "), vs[:, 2]) if len(v) else []
    fs = f.astype(str)
    body_f = np.char.add(np.char.add(np.char.add(np.char.add(np.char.add("f ", fs[:, 0]), " "), fs[:, 1]), " "), fs[:, 2]) if len(f) else []
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
        if len(v):
            fh.write("\n".join(body_v.tolist()) + "\n")
        if len(f):
            fh.write("\n".join(body_f.tolist()) + "\n")
    return True


class io:   # noqa: N801  (open3d.io)
    @staticmethod
    def write_triangle_mesh(filename, mesh, *a, **k):
        return write_obj_o3d(filename, mesh.vertices, mesh.triangles)


def read_obj(path, device="cpu"):
    """minimal Wavefront reader (v / f records, 1-based, `a/b/c` tolerated); parsed by the C library (csrc/objio.cu)"""
    import ctypes
    import numpy as np
    from . import _lib
    lib = _lib.load()
    nv, nf = ctypes.c_int64(0), ctypes.c_int64(0)
    _lib.check(lib.surfd_obj_read(os.fsencode(path), ctypes.byref(nv), ctypes.byref(nf), None, None))
    v = np.empty((nv.value, 3), dtype=np.float64)
    f = np.empty((nf.value, 3), dtype=np.int64)
    _lib.check(lib.surfd_obj_read(os.fsencode(path), ctypes.byref(nv), ctypes.byref(nf), ctypes.c_void_p(v.ctypes.data), ctypes.c_void_p(f.ctypes.data)))
    return torch.from_numpy(v).to(device), torch.from_numpy(f).to(device)


def _read_obj_py(path, device="cpu"):
    """the same reader in Python: the specification the native reader is tested against (tests only)"""
    import numpy as np
    vs, fs = [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("v "):
                vs.append(line.split()[1:4])
            elif line.startswith("f "):
                fs.append([p.split("/")[0] for p in line.split()[1:4]])
    v = torch.from_numpy(np.array(vs, dtype=np.float64).reshape(-1, 3))
    f = torch.from_numpy(np.array(fs, dtype=np.int64).reshape(-1, 3) - 1)
    return v.to(device), f.to(device)


# ---- filters ------------------------------------------------------------------------------------------------------
def _edge_groups(faces, n_verts):
    """per directed face edge (face-major (0,1) (1,2) (2,0)): multiplicity of its undirected edge; plus sorted keys"""
    e = faces[:, [0, 1, 1, 2, 2, 0]].reshape(-1, 2)
    es = torch.sort(e, dim=1).values
    key = es[:, 0] * n_verts + es[:, 1]
    _, inv, cnt = torch.unique(key, return_inverse=True, return_counts=True)
    return cnt[inv].reshape(-1, 3), key


def laplacian_smooth(vertices, faces, stepsmoothnum=3, boundary=True, cotangentweight=True):
    """MeshLab "Laplacian Smooth" (apply_coord_laplacian_smoothing defaults) -- float32 coordinates like MeshLab's CMeshO"""
    v = vertices.to(torch.float32).clone()
    f = faces.to(torch.int64)
    nv = v.shape[0]
    if f.shape[0] == 0:
        return v
    mult, _ = _edge_groups(f, nv)
    border_e = mult == 1                                         # [F,3]
    border_v = torch.zeros(nv, dtype=torch.bool, device=v.device)
    for j in range(3):
        sel = border_e[:, j]
        border_v[f[sel, j]] = True
        border_v[f[sel, (j + 1) % 3]] = True
    for _ in range(stepsmoothnum):
        acc = torch.zeros(nv, 3, dtype=torch.float32, device=v.device)
        wsum = torch.zeros(nv, dtype=torch.float32, device=v.device)
        for j in range(3):
            a, b, c = f[:, j], f[:, (j + 1) % 3], f[:, (j + 2) % 3]
            if cotangentweight:
                e1, e2 = v[a] - v[c], v[b] - v[c]
                den = torch.clamp(torch.linalg.norm(e1, dim=1) * torch.linalg.norm(e2, dim=1), min=1e-30)
                ang = torch.acos(torch.clamp((e1 * e2).sum(1) / den, -1.0, 1.0))
                w = torch.tan(torch.tensor(3.14159265358979323846 * 0.5, dtype=torch.float32, device=v.device) - ang)
            else:
                w = torch.ones(f.shape[0], dtype=torch.float32, device=v.device)
            sel = ~border_e[:, j]
            a_, b_, w_ = a[sel], b[sel], w[sel]
            acc.index_add_(0, a_, v[b_] * w_[:, None]); wsum.index_add_(0, a_, w_)
            acc.index_add_(0, b_, v[a_] * w_[:, None]); wsum.index_add_(0, b_, w_)
        acc[border_v] = 0
        wsum[border_v] = 0
        for j in range(3):
            sel = border_e[:, j]
            a_, b_ = f[sel, j], f[sel, (j + 1) % 3]
            one = torch.ones(a_.shape[0], dtype=torch.float32, device=v.device)
            acc.index_add_(0, a_, v[b_]); wsum.index_add_(0, a_, one)
            acc.index_add_(0, b_, v[a_]); wsum.index_add_(0, b_, one)
        ok = wsum > 0
        new = torch.where(ok[:, None], acc / torch.where(ok, wsum, torch.ones_like(wsum))[:, None], v)
        if not boundary:
            new = torch.where(border_v[:, None], v, new)
        v = new
    return v


def face_components(faces, n_verts):
    """component label (smallest face index of the component) per face; faces are adjacent when they share an edge"""
    f = faces.to(torch.int64)
    F = f.shape[0]
    dev = f.device
    parent = torch.arange(F, device=dev)
    if F == 0:
        return parent
    _, key = _edge_groups(f, n_verts)
    order = torch.argsort(key, stable=True)
    ks = key[order]
    face_of = order // 3
    same = ks[1:] == ks[:-1]
    fa, fb = face_of[:-1][same], face_of[1:][same]            # consecutive members of an edge group: a chain through the group
    while True:
        pa, pb = parent[fa], parent[fb]
        lo, hi = torch.minimum(pa, pb), torch.maximum(pa, pb)
        before = parent.clone()
        parent.scatter_reduce_(0, hi, lo, reduce="amin")
        while True:                                              # pointer jumping
            nxt = parent[parent]
            if torch.equal(nxt, parent):
                break
            parent = nxt
        if torch.equal(parent, before):
            break
    return parent


def remove_small_components(vertices, faces, mincomponentsize=2500, removeunref=True):
    """meshing_remove_connected_component_by_face_number: components with fewer faces than the threshold are deleted"""
    f = faces.to(torch.int64)
    lab = face_components(f, vertices.shape[0])
    _, inv, cnt = torch.unique(lab, return_inverse=True, return_counts=True)
    f = f[cnt[inv] >= mincomponentsize]
    if not removeunref:
        return vertices, f
    used = torch.zeros(vertices.shape[0], dtype=torch.bool, device=vertices.device)
    used[f.reshape(-1)] = True
    remap = torch.cumsum(used.to(torch.int64), 0) - 1
    return vertices[used], remap[f]


def write_obj_meshlab(path, vertices, faces):
    """MeshLab 2023.12 OBJ exporter layout (vcglib wrap/io_trimesh/export_obj.h) without normals / colours / texture"""
    import numpy as np
    v = vertices.detach().to(torch.float32).cpu().numpy().astype(np.float64)
    f = faces.detach().cpu().numpy()
    header = "####\n#\n# OBJ File Generated by Meshlab\n#\n####\n# Object %s\n#\n# Vertices: %d\n# Faces: %d\n#\n####\n" % (os.path.basename(path), len(v), len(f))
    _native_write(path, header, v, 0, "# %d vertices, 0 vertices normals\n\n" % len(v), f, "# %d faces, 0 coords texture\n\n# End of File\n" % len(f))


def _write_obj_meshlab_py(path, vertices, faces):
    """the same layout formatted by numpy / Python: the specification the native writer is tested against (tests only)"""
    import numpy as np
    v = vertices.detach().to(torch.float32).cpu().numpy()
    f = faces.detach().cpu().numpy().astype(np.int64) + 1
    name = os.path.basename(path)
    with open(path, "w") as fh:
        fh.write("####\n#\n# OBJ File Generated by Meshlab\n#\n####\n# Object %s\n#\n# Vertices: %d\n# Faces: %d\n#\n####\n" % (name, len(v), len(f)))
        if len(v):
            vs = np.char.mod("%f", v.astype(np.float64))
            fh.write("\n".join(np.char.add(np.char.add(np.char.add(np.char.add(np.char.add("v ", vs[:, 0]), " "), vs[:, 1]), " "), vs[:, 2]).tolist()) + "\n")
        fh.write("# %d vertices, 0 vertices normals\n\n" % len(v))
        if len(f):
            fs = f.astype(str)
            fh.write("\n".join(np.char.add(np.char.add(np.char.add(np.char.add(np.char.add("f ", fs[:, 0]), " "), fs[:, 1]), " "), fs[:, 2]).tolist()) + "\n")
        fh.write("# %d faces, 0 coords texture\n\n# End of File\n" % len(f))


class MeshSet:
    """the four pymeshlab.MeshSet calls of the scripts; the current mesh lives on `device`"""

    def __init__(self, device=None):
        self.device = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.vertices = None
        self.faces = None

    def set_verbosity(self, flag):
        pass

    def load_new_mesh(self, path):
        self.vertices, self.faces = read_obj(path, self.device)

    def set_current_mesh(self, vertices, faces):
        """device-resident hand-over (skips the .obj round trip of the scripts)"""
        self.vertices, self.faces = vertices.to(self.device), faces.to(self.device)

    def apply_coord_laplacian_smoothing(self, stepsmoothnum=3, boundary=True, cotangentweight=True, selected=False):
        if selected:
            raise NotImplementedError("selected=True")
        self.vertices = laplacian_smooth(self.vertices, self.faces, stepsmoothnum, boundary, cotangentweight)

    def meshing_remove_connected_component_by_face_number(self, mincomponentsize=25, removeunref=True):
        self.vertices, self.faces = remove_small_components(self.vertices, self.faces, mincomponentsize, removeunref)

    def save_current_mesh(self, path):
        write_obj_meshlab(path, self.vertices, self.faces)


def finish_and_save(vertices, faces, mesh_path, mincomponentsize=2500):
    """generate_uncond.py:113-122 in one call, device-resident: smoothing + small-component removal + one file write"""
    v = laplacian_smooth(vertices, faces)
    v, f = remove_small_components(v, faces, mincomponentsize)
    d = os.path.dirname(mesh_path)
    if d:
        os.makedirs(d, exist_ok=True)
    write_obj_meshlab(mesh_path, v, f)
    return v, f
