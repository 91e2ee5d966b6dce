"""CPU restatement (numpy) of the rectify / resize / colour->gray pre-stage (b200AprilTagsSetRectification) -- TEST
INFRASTRUCTURE ONLY, like the rest of oracle/.

The map is OpenCV's initUndistortRectifyMap (imgproc/src/undistort.dispatch.cpp: [x y w] = (P R)^-1 [u v 1], plumb_bob /
rational distortion, source = K * distorted point), evaluated per pixel in double in exactly the operation order of
isaac_ros_apriltag_b200/csrc/capi.cu (b200AprilTagsSetRectification) and stored as float32 like OpenCV's CV_32FC1 maps; the remap is
bilinear on the gray values of the four source pixels in float32, constant 0 outside the source, rounded half up -- the order of
k_rectify (csrc/k_dense.cu).  tests/test_rectify.py pins it against cv2.initUndistortRectifyMap + cv2.remap (+-1 gray level: OpenCV
interpolates with 1/32-pixel fixed-point weights)."""
import numpy as np

from . import oracle as O


def rectify_map(K, D, R, P, width, height):
    K, R, P = (np.asarray(m, np.float64).reshape(3, 3) for m in (K, R, P))
    D = np.concatenate([np.asarray(D, np.float64).reshape(-1), np.zeros(8)])[:8]
    A = np.empty((3, 3))
    for i in range(3):
        for j in range(3):
            A[i, j] = P[i, 0] * R[0, j] + P[i, 1] * R[1, j] + P[i, 2] * R[2, j]
    a = A.reshape(-1)
    det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6])
    idet = 1.0 / det
    iR = np.array([(a[4] * a[8] - a[5] * a[7]) * idet, (a[2] * a[7] - a[1] * a[8]) * idet, (a[1] * a[5] - a[2] * a[4]) * idet,
                   (a[5] * a[6] - a[3] * a[8]) * idet, (a[0] * a[8] - a[2] * a[6]) * idet, (a[2] * a[3] - a[0] * a[5]) * idet,
                   (a[3] * a[7] - a[4] * a[6]) * idet, (a[1] * a[6] - a[0] * a[7]) * idet, (a[0] * a[4] - a[1] * a[3]) * idet])
    fx, fy, u0, v0 = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    k1, k2, p1, p2, k3, k4, k5, k6 = D
    u = np.arange(width, dtype=np.float64)[None, :]
    v = np.arange(height, dtype=np.float64)[:, None]
    X = u * iR[0] + v * iR[1] + iR[2]
    Y = u * iR[3] + v * iR[4] + iR[5]
    Wq = u * iR[6] + v * iR[7] + iR[8]
    w = 1.0 / Wq
    x = X * w
    y = Y * w
    x2 = x * x
    y2 = y * y
    r2 = x2 + y2
    xy2 = 2 * x * y
    kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / (1 + ((k6 * r2 + k5) * r2 + k4) * r2)
    xd = x * kr + p1 * xy2 + p2 * (r2 + 2 * x2)
    yd = y * kr + p1 * (r2 + 2 * y2) + p2 * xy2
    return (fx * xd + u0).astype(np.float32), (fy * yd + v0).astype(np.float32)


def rectify_gray(raw, encoding, mapx, mapy):
    """raw: (h, w[, c]) uint8 in `encoding`; returns the rectified gray image (mapx.shape) uint8."""
    gray = raw if encoding == "mono8" else O.to_gray(raw, encoding)
    sh, sw = gray.shape
    g = np.zeros((sh + 2, sw + 2), np.float32)  # constant 0 border
    g[1:-1, 1:-1] = gray
    inside = (mapx > np.float32(-1)) & (mapy > np.float32(-1)) & (mapx < np.float32(sw)) & (mapy < np.float32(sh))
    mx = np.where(inside, mapx, np.float32(0))
    my = np.where(inside, mapy, np.float32(0))
    fx0 = np.floor(mx)
    fy0 = np.floor(my)
    ax = (mx - fx0).astype(np.float32)
    ay = (my - fy0).astype(np.float32)
    x0 = fx0.astype(np.int64) + 1
    y0 = fy0.astype(np.int64) + 1
    one = np.float32(1)
    g00, g01, g10, g11 = g[y0, x0], g[y0, x0 + 1], g[y0 + 1, x0], g[y0 + 1, x0 + 1]
    top = (g00 * (one - ax)).astype(np.float32) + (g01 * ax).astype(np.float32)
    bot = (g10 * (one - ax)).astype(np.float32) + (g11 * ax).astype(np.float32)
    v = (top * (one - ay)).astype(np.float32) + (bot * ay).astype(np.float32)
    v = np.where(inside, v, np.float32(0))
    return (v + np.float32(0.5)).astype(np.int32).astype(np.uint8)
