"""Host logic of the plugin mirror (no GPU): the reference's three gtests
(/root/reference/isaac_ros_apriltag/test/apriltag_node_test.cpp:29-89) re-expressed against the ROS-free node core,
plus the marshalling helpers."""
import ctypes as C

import numpy as np
import pytest

from isaac_ros_apriltag_b200 import node


def test_invalid_tag_family():  # apriltag_node_test.cpp:29-49
    with pytest.raises(RuntimeError, match="Tag family not supported by specified backend"):
        node.AprilTagNode(tag_family="NOTHING")


def test_unsupported_tag_family():  # apriltag_node_test.cpp:51-72  (tag36h10 on the default CUDA backend)
    with pytest.raises(RuntimeError, match="Tag family not supported by specified backend"):
        node.AprilTagNode(tag_family="tag36h10", backends="CUDA")


def test_supported_tag_family():  # apriltag_node_test.cpp:74-89
    n = node.AprilTagNode(tag_family="tag36h10", backends="CPU")
    assert not n.using_cuapriltag_impl()
    n.close()


def test_backend_selection_rule():  # apriltag_node.cpp:576-582: exactly CUDA -> cuAprilTag impl, anything else -> VPI impl
    L = node.lib()
    assert L.b200NodeParseBackends(b"CUDA") == 2
    assert L.b200NodeParseBackends(b"CPU,CUDA") == 3
    assert L.b200NodeParseBackends(b"PVA") == 4
    assert node.AprilTagNode().using_cuapriltag_impl()
    assert not node.AprilTagNode(backends="CPU,CUDA").using_cuapriltag_impl()
    # all nine family names are accepted by the non-CUDA strategy at construction (apriltag_node.cpp:182-191)
    for fam in ["tag36h11", "tag16h5", "tag25h9", "tag36h10", "circle21h7", "circle49h12", "custom48h12", "standard41h12",
                "standard52h13"]:
        node.AprilTagNode(tag_family=fam, backends="CPU").close()


def _quat(m, col_major, normalize):
    a = (C.c_float * 9)(*[float(v) for v in m])
    o = (C.c_double * 4)()
    node.lib().b200NodeRotationToQuaternion(a, int(col_major), int(normalize), o)
    return np.array(o[:])  # x y z w


def test_rotation_to_quaternion():
    rz180 = np.diag([-1.0, -1.0, 1.0])
    q = _quat(rz180.T.reshape(-1), True, False)  # column major input
    assert np.allclose(np.abs(q), [0, 0, 1, 0], atol=1e-6)  # POL golden orientation (w=0, z=1)
    assert np.allclose(_quat(np.eye(3).reshape(-1), False, True), [0, 0, 0, 1])
    rng = np.random.default_rng(3)
    for _ in range(20):
        A = rng.normal(size=(3, 3))
        U, _, Vt = np.linalg.svd(A)
        R = U @ Vt
        if np.linalg.det(R) < 0:
            R = -R
        x, y, z, w = _quat(R.reshape(-1), False, True)
        Rq = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                       [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                       [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert np.abs(Rq - R).max() < 1e-5
        assert np.allclose(_quat(R.T.reshape(-1), True, True), [x, y, z, w], atol=1e-6)
