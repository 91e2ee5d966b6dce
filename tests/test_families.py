"""Code tables: the committed generated files are what tools/gen_families.py produces from the OpenCV dictionaries, and the
three copies (oracle .inc, device .inc, Python JSON) agree."""
import importlib.util
import os

from isaac_ros_apriltag_b200.families import families, tag_cells

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gen():
    spec = importlib.util.spec_from_file_location("gen_families", os.path.join(ROOT, "tools", "gen_families.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_generated_tables_are_reproducible():
    m = _gen()
    text = m.emit()  # asserts the upstream table heads and the minimum inter-code distances while generating
    for rel in ("oracle/tag_families_data.inc", "isaac_ros_apriltag_b200/csrc/tag_families_data.inc"):
        assert open(os.path.join(ROOT, rel)).read() == text, rel


def test_json_matches_generator_and_layout():
    m = _gen()
    fams = families()
    for name, dict_id, d, h in m.FAMILIES:
        codes, bx, by = m.family_codes(dict_id, d)
        f = fams[name]
        assert f["codes"] == codes and f["bit_x"] == bx and f["bit_y"] == by
        assert f["nbits"] == d * d and f["width_at_border"] == d + 2 and f["total_width"] == d + 4 and f["h"] == h
        # spiral layout: a quarter turn of the cell grid is a cyclic shift of the bit string by nbits/4 (rotate90)
        q = (d * d) // 4
        for k in range(q):
            x, y = bx[k], by[k]
            assert (bx[k + q], by[k + q]) == (d + 1 - y, x)
    cells = tag_cells("tag36h11", 0)
    assert cells.shape == (10, 10) and cells[0].all() and not cells[1, 1:9].any()  # white ring, black border


def test_oracle_registered_reversed_border_family():
    """ato_register_family: a synthetic reversed-border family (tag36h11's table, white border on black) decodes only against itself
    -- the oracle side of tests/test_gpu_parity.py::test_registered_reversed_border_family."""
    import numpy as np
    from isaac_ros_apriltag_b200 import families as F, synth
    from oracle import oracle as O
    fam = dict(F.families()["tag36h11"])
    fam["reversed_border"] = True
    F.add_family("custom0", fam)
    O.register_family(4, fam)
    g, truth = synth.make_frame(np.random.default_rng(9), 960, 720, [("custom0", 5), ("custom0", 77), ("tag36h11", 5)], side_px=(100, 170),
                                max_tilt_deg=25)
    got = {fams: sorted((d["family"], d["id"]) for d in O.Oracle(fams).detect(g) if d["hamming"] == 0)
           for fams in (("custom0",), ("tag36h11",), ("tag36h11", "custom0"))}
    assert got[("custom0",)] == [("custom0", 5), ("custom0", 77)]
    assert got[("tag36h11",)] == [("tag36h11", 5)]
    assert got[("tag36h11", "custom0")] == [("custom0", 5), ("custom0", 77), ("tag36h11", 5)]
    for d in O.Oracle(("custom0",)).detect(g):
        tr = [t for t in truth if t["family"] == "custom0" and t["id"] == d["id"]][0]
        assert np.abs(d["p"] - tr["p"]).max() < 0.6
