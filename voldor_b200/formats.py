"""On-disk formats either side of the window solver.  numpy + zlib only (no OpenCV).

Behavioural sources (reference):
  * Middlebury `.flo`:  slam_py/flow_utils.py:10-25, voldor/utils.cpp:23-41 — float32 magic 202021.25,
    int32 width, int32 height, then H x W x 2 float32 (u, v) row-major, little endian.
  * disparity maps: slam_py/voldor_slam.py:302-307 — a `.flo` whose first channel is the negated disparity, or a
    16-bit grey PNG storing disparity * 256 (KITTI convention).
  * trajectory files: slam_py/voldor_slam.py:317-329 — KITTI (12 numbers of the 3x4 pose per line) and TartanAir
    (tz tx ty qz qx qy qw per line).
"""
import struct
import zlib

import numpy as np

FLO_MAGIC = np.float32(202021.25)


def load_flow(path):
    """H x W x 2 float32, or None when the magic number does not match (flow_utils.py:10-18)"""
    with open(path, "rb") as f:
        head = f.read(12)
        if len(head) < 12:
            return None
        magic, w, h = struct.unpack("<fii", head)
        if np.float32(magic) != FLO_MAGIC:
            return None
        data = np.frombuffer(f.read(h * w * 8), dtype="<f4")
    if data.size != h * w * 2:
        raise ValueError(f"{path}: truncated .flo ({data.size} of {h * w * 2} floats)")
    return data.reshape(h, w, 2).astype(np.float32)


def save_flow(path, flow):
    flow = np.ascontiguousarray(flow, dtype="<f4")
    h, w = flow.shape[:2]
    with open(path, "wb") as f:
        f.write(struct.pack("<fii", float(FLO_MAGIC), w, h))
        f.write(flow.tobytes())


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = np.abs(p - a), np.abs(p - b), np.abs(p - c)
    return np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c))


def read_png_gray(path):
    """non-interlaced 8/16-bit greyscale PNG -> uint8/uint16 H x W (the subset KITTI disparity files use)"""
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError(f"{path}: not a PNG file")
    pos, idat, hdr = 8, [], None
    while pos < len(raw):
        (n,), typ = struct.unpack(">I", raw[pos:pos + 4]), raw[pos + 4:pos + 8]
        body = raw[pos + 8:pos + 8 + n]
        if typ == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"IDAT":
            idat.append(body)
        elif typ == b"IEND":
            break
        pos += 12 + n
    w, h, depth, ctype, _, _, interlace = hdr
    if ctype != 0 or depth not in (8, 16) or interlace != 0:
        raise ValueError(f"{path}: only non-interlaced 8/16-bit greyscale PNG is supported")
    bpp = depth // 8
    stride = w * bpp
    data = np.frombuffer(zlib.decompress(b"".join(idat)), np.uint8).reshape(h, stride + 1)
    out = np.zeros((h, stride), np.int32)
    prev = np.zeros(stride, np.int32)
    for y in range(h):
        ft, line = int(data[y, 0]), data[y, 1:].astype(np.int32)
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        else:
            cur = np.zeros(stride, np.int32)
            for i in range(stride):  # filters 1, 3, 4 depend on the already decoded left neighbour
                a = cur[i - bpp] if i >= bpp else 0
                b = prev[i]
                c = prev[i - bpp] if i >= bpp else 0
                if ft == 1:
                    pred = a
                elif ft == 3:
                    pred = (a + b) >> 1
                elif ft == 4:
                    pred = int(_paeth(np.int32(a), np.int32(b), np.int32(c)))
                else:
                    raise ValueError(f"{path}: bad PNG filter {ft}")
                cur[i] = (line[i] + pred) & 255
        out[y] = cur
        prev = cur
    out = out.astype(np.uint8)
    if depth == 8:
        return out.reshape(h, w)
    return out.reshape(h, w, 2).astype(np.uint16)[..., 0] * 256 + out.reshape(h, w, 2)[..., 1]


def write_png_gray16(path, img):
    """uint16 H x W -> 16-bit greyscale PNG (filter 0); used by tests and for exporting disparity"""
    img = np.ascontiguousarray(img, dtype=">u2")
    h, w = img.shape
    rows = np.concatenate([np.zeros((h, 1), np.uint8), img.view(np.uint8).reshape(h, w * 2)], axis=1)

    def chunk(typ, body):
        return struct.pack(">I", len(body)) + typ + body + struct.pack(">I", zlib.crc32(typ + body) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 16, 0, 0, 0, 0)))
        f.write(chunk(b"IDAT", zlib.compress(rows.tobytes(), 6)))
        f.write(chunk(b"IEND", b""))


def load_disparity(path):
    """float32 H x W disparity in pixels (voldor_slam.py:302-307)"""
    if path.endswith(".flo"):
        return np.ascontiguousarray(-load_flow(path)[..., 0])
    if path.endswith(".png"):
        return read_png_gray(path).astype(np.float32) / 256.0
    raise ValueError(f"Unsupported disparity format {path}")


def _quat_xyzw(R):
    """rotation matrix -> unit quaternion (x, y, z, w), w >= 0 branch selection as in Shepperd's method"""
    R = np.asarray(R, np.float64)
    k = np.array([R[0, 0] + R[1, 1] + R[2, 2], R[0, 0], R[1, 1], R[2, 2]])
    i = int(np.argmax(k))
    if i == 0:
        w = np.sqrt(1 + k[0]) / 2
        q = np.array([(R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w), w])
    else:
        a, b, c = i - 1, i % 3, (i + 1) % 3
        s = np.sqrt(1 + R[a, a] - R[b, b] - R[c, c]) * 2
        q = np.zeros(4)
        q[a] = s / 4
        q[b] = (R[b, a] + R[a, b]) / s
        q[c] = (R[c, a] + R[a, c]) / s
        q[3] = (R[c, b] - R[b, c]) / s
    return q / np.linalg.norm(q)


def save_poses(path, Tcw_list, format="KITTI"):
    """one line per frame (voldor_slam.py:317-329); Tcw_list: sequence of 4x4 (or 3x4) camera-to-world poses"""
    with open(path, "w") as f:
        for T in Tcw_list:
            T = np.asarray(T)
            if format == "KITTI":
                f.write(" ".join(str(v) for v in T[:3, :4].reshape(-1)) + "\n")
            elif format == "TartanAir":
                q, t = _quat_xyzw(T[:3, :3]), T[:3, 3]
                f.write(f"{t[2]} {t[0]} {t[1]} {q[2]} {q[0]} {q[1]} {q[3]}\n")
            else:
                raise ValueError(f"Unsupported pose format {format}")


def load_poses_kitti(path):
    """N x 4 x 4 poses from a KITTI trajectory file"""
    rows = np.loadtxt(path, dtype=np.float64).reshape(-1, 3, 4)
    T = np.tile(np.eye(4), (rows.shape[0], 1, 1))
    T[:, :3, :4] = rows
    return T


def accumulate_poses(poses, Twc0=None):
    """chain window poses (rvec, t of frame i-1 -> i, as returned by voldor()) into camera-to-world matrices;
    voldor_slam.py:518 left-multiplies its world-to-current transform by T(rvec, t) and stores the inverse in each
    frame (`Frame(np.linalg.inv(self.Twc_cur))`, :509), which is what save_poses() writes."""
    T = np.eye(4) if Twc0 is None else np.array(Twc0, np.float64)
    out = [T.copy()]
    for p in np.asarray(poses, np.float64):
        th = np.linalg.norm(p[:3])
        K = np.zeros((3, 3))
        if th > 1e-12:
            k = p[:3] / th
            K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
        M = np.eye(4)
        M[:3, :3], M[:3, 3] = R, p[3:6]
        T = T @ np.linalg.inv(M)
        out.append(T.copy())
    return out
