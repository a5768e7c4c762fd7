"""TEST INFRASTRUCTURE (never imported by the product): numpy / networkx restatement of the mesh clean-up that
get_mesh_from_udf performs through trimesh (meshudf/meshudf.py:379-434) and of the output stage that the scripts perform
through pymeshlab (sample/generate_uncond.py:113-122).

PARITY UNPINNED: trimesh 4.0.8, pymeshlab 2023.12 and open3d 0.18.0 are third-party dependencies of the reference
(environment.yaml:105-239) that are absent from this image (no network).  Each function below restates the PUBLISHED
algorithm of the named library function from its documentation / source as of that version; nothing here was validated
against the libraries themselves.  The reference's own arithmetic on this path (border detection, the neighbour lists and
the 20-iteration lambda=0.3 Laplacian of meshudf.py:404-434) is restated literally.
"""
from collections import defaultdict

import numpy as np

TOL_MERGE = 1e-8


# ---- trimesh.grouping -------------------------------------------------------------------------------------------
def unique_rows(data, keep_order=False):
    """grouping.unique_rows: (indices of the first occurrence of every distinct row, inverse)"""
    data = np.asanyarray(data)
    _, unique, inverse = np.unique(data, axis=0, return_index=True, return_inverse=True)
    inverse = inverse.reshape(-1)
    if keep_order:
        order = np.argsort(unique)
        rank = np.empty_like(order)
        rank[order] = np.arange(len(order))
        return unique[order], rank[inverse]
    return unique, inverse


def merge_vertices(vertices, faces):
    """grouping.merge_vertices (no uv / normals): round(vertices * 1e8) rows, referenced vertices only, first occurrence order"""
    vertices = np.asanyarray(vertices, dtype=np.float64)
    referenced = np.zeros(len(vertices), dtype=bool)
    referenced[faces] = True
    stacked = (vertices * (10 ** 8)).round().astype(np.int64)
    u, i = unique_rows(stacked[referenced], keep_order=True)
    inverse = np.zeros(len(vertices), dtype=np.int64)
    inverse[referenced] = i
    mask = np.nonzero(referenced)[0][u]
    return vertices[mask], inverse[faces]


def unique_faces(faces):
    mask = np.zeros(len(faces), dtype=bool)
    mask[unique_rows(np.sort(faces, axis=1))[0]] = True
    return mask


def nondegenerate(vertices, faces, height=TOL_MERGE):
    """triangles.nondegenerate via triangles.extents: (longest edge, 2 * area / longest edge) both > height"""
    tri = vertices[faces]
    edges = tri[:, [0, 1, 2]] - tri[:, [1, 2, 0]]
    length = np.sqrt((edges ** 2).sum(axis=2))
    base = length.max(axis=1)
    cross = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    area = np.sqrt((cross ** 2).sum(axis=1)) * 0.5
    with np.errstate(divide="ignore", invalid="ignore"):
        h = (area * 2) / base
    return (base > height) & (h > height)


def process(vertices, faces):
    """mesh.process(validate=False); remove_duplicate_faces(); remove_degenerate_faces()"""
    vertices, faces = merge_vertices(vertices, faces)
    faces = faces[unique_faces(faces)]
    faces = faces[nondegenerate(vertices, faces)]
    return vertices, faces


def faces_to_edges(faces):
    return np.asanyarray(faces)[:, [0, 1, 1, 2, 2, 0]].reshape(-1, 2)


def group_rows_count1(rows):
    """grouping.group_rows(rows, require_count=1): indices of rows that occur exactly once"""
    _, inverse, counts = np.unique(rows, axis=0, return_inverse=True, return_counts=True)
    return np.nonzero(counts[inverse.reshape(-1)] == 1)[0]


def fill_holes(vertices, faces):
    """repair.fill_holes: cycles of the boundary-edge graph (networkx.cycle_basis) with 3 or 4 vertices get 1 or 2 faces,
    wound against the mesh edge they share"""
    import networkx as nx
    if len(faces) < 3:
        return faces
    edges = faces_to_edges(faces)
    edges_sorted = np.sort(edges, axis=1)
    boundary_groups = group_rows_count1(edges_sorted)
    if len(boundary_groups) < 3:
        return faces
    boundary_edges = edges[boundary_groups]
    g = nx.Graph()
    for (a, b), i in zip(boundary_edges, boundary_groups):
        g.add_edge(int(a), int(b), index=int(i))
    new_faces = []
    for hole in nx.cycle_basis(g):
        hole = np.asanyarray(hole)
        if len(hole) == 3:
            new_faces.append(hole)
        elif len(hole) == 4:
            new_faces.append(hole[[0, 1, 2]])
            new_faces.append(hole[[2, 3, 0]])
    if not new_faces:
        return faces
    new_faces = np.array(new_faces, dtype=np.int64)
    for fi, face in enumerate(new_faces):
        edge_test = face[:2]
        data = g.get_edge_data(int(edge_test[0]), int(edge_test[1]))
        if data is None:     # a quad's second triangle starts on the diagonal: test its boundary edge instead
            edge_test = face[1:]
            data = g.get_edge_data(int(edge_test[0]), int(edge_test[1]))
        edge_boundary = edges[data["index"]]
        if not (edge_test[0] == edge_boundary[1]):
            new_faces[fi] = face[::-1]
    return np.vstack((faces, new_faces))


def smooth_borders(vertices, faces, iterations=20, lambda_=0.3):
    """meshudf.py:404-434, literally (dict of neighbour lists, dense row average instead of the scipy coo matrix)"""
    vertices = np.array(vertices, dtype=np.float64)
    edges_sorted = np.sort(faces_to_edges(faces), axis=1)
    border_edges = group_rows_count1(edges_sorted)
    neighbours = defaultdict(lambda: [])
    for u, v in edges_sorted[border_edges]:
        neighbours[int(u)].append(int(v))
        neighbours[int(v)].append(int(u))
    border_vertices = np.array(list(neighbours.keys()), dtype=np.int64)
    if len(border_vertices) == 0:
        return vertices
    for _ in range(iterations):
        avg = np.stack([vertices[ns].sum(axis=0) / len(ns) for ns in neighbours.values()])
        laplacian = avg - vertices[border_vertices]
        vertices[border_vertices] = vertices[border_vertices] + lambda_ * laplacian
    return vertices


def clean_mesh(vertices, faces, smooth=True):
    """meshudf.py:379-437 from the filtered faces to (float32 vertices, int64 faces)"""
    vertices = np.asanyarray(vertices, dtype=np.float64)
    faces = np.asanyarray(faces, dtype=np.int64)
    vertices, faces = merge_vertices(vertices, faces)
    vertices, faces = process(vertices, faces)
    faces = fill_holes(vertices, faces)
    vertices, faces = merge_vertices(vertices, faces)
    n_verts, n_faces, n_iter = 0, 0, 0
    while (n_verts, n_faces) != (len(vertices), len(faces)) and n_iter < 10:
        vertices, faces = process(vertices, faces)
        n_verts, n_faces = len(vertices), len(faces)
        n_iter += 1
        vertices, faces = merge_vertices(vertices, faces)
    vertices, faces = merge_vertices(vertices, faces)
    if smooth and len(faces):
        vertices = smooth_borders(vertices, faces)
    return vertices.astype(np.float32), faces


# ---- output stage (pymeshlab 2023.12 filters used by sample/generate_uncond.py:113-122) -----------------------------
def laplacian_smooth(vertices, faces, stepsmoothnum=3, boundary=True, cotangentweight=True):
    """`apply_coord_laplacian_smoothing()` with its defaults (stepsmoothnum=3, boundary=True, cotangentweight=True,
    selected=False): MeshLab's "Laplacian Smooth" = vcg::tri::Smooth::VertexCoordLaplacian(m, step, SmoothSelected=false,
    cotangentFlag): per step every vertex moves to the weighted mean of its edge neighbours (Jacobi), accumulating per
    FACE edge (interior edges are therefore visited from both faces); border edges are then re-accumulated alone so that
    border vertices only follow the border polyline ("1D boundary smoothing").  Cotangent weights as in vcglib:
    w = tan(pi/2 - angle opposite to the edge) per face."""
    v = np.array(vertices, dtype=np.float32)     # MeshLab's CMeshO stores float32 coordinates
    f = np.asanyarray(faces, dtype=np.int64)
    es = np.sort(faces_to_edges(f), axis=1)
    _, inv, cnt = np.unique(es, axis=0, return_inverse=True, return_counts=True)
    is_border_edge = (cnt[inv.reshape(-1)] == 1).reshape(-1, 3)           # [F, 3] for edges (0,1) (1,2) (2,0)
    border_vertex = np.zeros(len(v), dtype=bool)
    for j in range(3):
        sel = is_border_edge[:, j]
        border_vertex[f[sel, j]] = True
        border_vertex[f[sel, (j + 1) % 3]] = True
    for _ in range(stepsmoothnum):
        acc = np.zeros((len(v), 3), dtype=np.float32)
        wsum = np.zeros(len(v), dtype=np.float32)
        for j in range(3):
            a, b, c = f[:, j], f[:, (j + 1) % 3], f[:, (j + 2) % 3]
            interior = ~is_border_edge[:, j]
            if cotangentweight:
                e1 = v[a] - v[c]
                e2 = v[b] - v[c]
                cosang = (e1 * e2).sum(1) / np.maximum(np.linalg.norm(e1, axis=1) * np.linalg.norm(e2, axis=1), 1e-30)
                ang = np.arccos(np.clip(cosang, -1.0, 1.0))
                w = np.tan(np.float32(np.pi * 0.5) - ang).astype(np.float32)
            else:
                w = np.ones(len(f), dtype=np.float32)
            sel = interior
            np.add.at(acc, a[sel], v[b[sel]] * w[sel, None]); np.add.at(wsum, a[sel], w[sel])
            np.add.at(acc, b[sel], v[a[sel]] * w[sel, None]); np.add.at(wsum, b[sel], w[sel])
        # border vertices: reset, then accumulate border edges only (weight 1)
        acc[border_vertex] = 0
        wsum[border_vertex] = 0
        for j in range(3):
            sel = is_border_edge[:, j]
            a, b = f[sel, j], f[sel, (j + 1) % 3]
            np.add.at(acc, a, v[b]); np.add.at(wsum, a, 1.0)
            np.add.at(acc, b, v[a]); np.add.at(wsum, b, 1.0)
        ok = wsum > 0
        new = v.copy()
        new[ok] = acc[ok] / wsum[ok, None]
        if not boundary:
            new[border_vertex] = v[border_vertex]
        v = new
    return v


def face_components(faces, n_verts):
    """connected components over face-face adjacency through shared EDGES (vcg::tri::Clean::ConnectedComponents uses FF
    topology): label per face"""
    f = np.asanyarray(faces, dtype=np.int64)
    parent = np.arange(len(f))

    def find(i):
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    es = np.sort(faces_to_edges(f), axis=1)
    key = es[:, 0] * n_verts + es[:, 1]
    order = np.argsort(key, kind="stable")
    ks = key[order]
    face_of = order // 3
    start = 0
    for i in range(1, len(ks) + 1):
        if i == len(ks) or ks[i] != ks[start]:
            for j in range(start + 1, i):
                ra, rb = find(face_of[start]), find(face_of[j])
                if ra != rb:
                    parent[max(ra, rb)] = min(ra, rb)
            start = i
    return np.array([find(i) for i in range(len(f))])


def remove_small_components(vertices, faces, mincomponentsize=2500, removeunref=True):
    """`meshing_remove_connected_component_by_face_number(mincomponentsize=2500)` (removeunref defaults to True):
    vcg::tri::Clean::RemoveSmallConnectedComponentsSize deletes components with FEWER faces than the threshold, then
    unreferenced vertices are removed (vertex order kept)."""
    f = np.asanyarray(faces, dtype=np.int64)
    lab = face_components(f, len(vertices))
    _, inv, cnt = np.unique(lab, return_inverse=True, return_counts=True)
    keep = cnt[inv.reshape(-1)] >= mincomponentsize
    f = f[keep]
    if not removeunref:
        return np.asanyarray(vertices), f
    used = np.zeros(len(vertices), dtype=bool)
    used[f] = True
    remap = np.cumsum(used) - 1
    return np.asanyarray(vertices)[used], remap[f]
