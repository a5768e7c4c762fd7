"""Host-side .obj text conversion in the C library (csrc/objio.cu: surfd_obj_write / surfd_obj_read) against the numpy / Python
formatting it replaces (surfd_b200/output.py `_write_obj_*_py`, `_read_obj_py`): byte-identical files, identical arrays.  No GPU."""
import os

import numpy as np
import pytest
import torch

from surfd_b200 import output as O


def _same_file(a, b):
    da = open(a, "rb").read().replace(os.path.basename(a).encode(), b"NAME").replace(os.path.splitext(os.path.basename(a))[0].encode(), b"STEM")
    db = open(b, "rb").read().replace(os.path.basename(b).encode(), b"NAME").replace(os.path.splitext(os.path.basename(b))[0].encode(), b"STEM")
    return da == db


@pytest.mark.parametrize("seed", [0, 1])
def test_writers_are_byte_identical_to_the_python_formatting(tmp_path, seed):
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(4000, 3, generator=g, dtype=torch.float64) * torch.tensor([1e-7, 1.0, 3e6], dtype=torch.float64)
    v[0] = torch.tensor([0.0, -0.0, 1e-5]); v[1] = torch.tensor([123456.5, 1e6, -1e-4]); v[2] = torch.tensor([999999.5, 0.1 + 0.2, -2.5e-7])
    v[3] = torch.tensor([1e15, -1e22, 5e-324]); v[4] = torch.tensor([0.5, 1.5, 2.5])
    f = torch.randint(0, 4000, (7000, 3), generator=g)
    for native, spec, tag in ((O.write_obj_meshlab, O._write_obj_meshlab_py, "ml"), (O.write_obj_o3d, O._write_obj_o3d_py, "o3")):
        a, b = str(tmp_path / f"mesha{tag}.obj"), str(tmp_path / f"meshb{tag}.obj")
        native(a, v, f)
        spec(b, v, f)
        assert _same_file(a, b), tag
        rv, rf = O.read_obj(a)
        pv, pf = O._read_obj_py(a)
        assert torch.equal(rv, pv) and torch.equal(rf, pf) and torch.equal(rf, f)
    # empty mesh (pymeshlab writes one after removing every component)
    e = torch.zeros(0, 3)
    a, b = str(tmp_path / "emptya.obj"), str(tmp_path / "emptyb.obj")
    O.write_obj_meshlab(a, e, e.long()); O._write_obj_meshlab_py(b, e, e.long())
    assert _same_file(a, b)
    rv, rf = O.read_obj(a)
    assert tuple(rv.shape) == (0, 3) and tuple(rf.shape) == (0, 3)


def test_reader_skips_what_it_does_not_know_and_reports_errors(tmp_path):
    p = str(tmp_path / "m.obj")
    with open(p, "w") as fh:
        fh.write("# comment\nmtllib x.mtl\nv 1 2 3\nv  -0.5 1e-3 4.25 0.1 0.2 0.3\nvn 0 0 1\nvt 0 0\nv 7 8 9\nf 1/1/1 2/2/1 3/3/1\nf 3 2 1\ng grp\nf 1//1 3//1 2//1")
    v, f = O.read_obj(p)
    pv, pf = O._read_obj_py(p)
    assert torch.equal(v, pv) and torch.equal(f, pf)
    assert v.tolist() == [[1, 2, 3], [-0.5, 1e-3, 4.25], [7, 8, 9]] and f.tolist() == [[0, 1, 2], [2, 1, 0], [0, 2, 1]]
    with pytest.raises(Exception):
        O.read_obj(str(tmp_path / "missing.obj"))
    with open(p, "w") as fh:
        fh.write("v 1 2\n")
    with pytest.raises(Exception):
        O.read_obj(p)
    with pytest.raises(Exception):
        O.write_obj_meshlab(str(tmp_path / "no_such_dir" / "x.obj"), torch.zeros(1, 3), torch.zeros(1, 3, dtype=torch.int64))
