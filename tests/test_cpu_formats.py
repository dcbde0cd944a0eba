"""File formats either side of the solver: .flo, 16-bit disparity PNG, KITTI / TartanAir poses."""
import struct
import zlib

import numpy as np
import pytest

import synth
from voldor_b200 import formats


def test_flo_round_trip_and_layout(tmp_path):
    rng = np.random.default_rng(0)
    flow = rng.normal(size=(7, 11, 2)).astype(np.float32)
    p = str(tmp_path / "a.flo")
    formats.save_flow(p, flow)
    raw = open(p, "rb").read()
    # reference layout (flow_utils.py:10-25): magic, int32 w, int32 h, data
    assert struct.unpack("<f", raw[:4])[0] == 202021.25
    assert struct.unpack("<ii", raw[4:12]) == (11, 7)
    assert len(raw) == 12 + 7 * 11 * 8
    assert np.array_equal(formats.load_flow(p), flow)


def test_flo_bad_magic_and_truncation(tmp_path):
    p = str(tmp_path / "bad.flo")
    open(p, "wb").write(struct.pack("<fii", 1.0, 2, 2) + b"\0" * 32)
    assert formats.load_flow(p) is None
    open(p, "wb").write(struct.pack("<fii", 202021.25, 4, 4) + b"\0" * 8)
    with pytest.raises(ValueError):
        formats.load_flow(p)


def _png_with_filters(path, img16):
    """16-bit grey PNG whose rows cycle through all five PNG filter types"""
    h, w = img16.shape
    be = np.ascontiguousarray(img16, ">u2").view(np.uint8).reshape(h, w * 2).astype(np.int32)
    rows = bytearray()
    prev = np.zeros(w * 2, np.int32)
    for y in range(h):
        ft = y % 5
        cur = be[y]
        a = np.concatenate([np.zeros(2, np.int32), cur[:-2]])
        c = np.concatenate([np.zeros(2, np.int32), prev[:-2]])
        pred = [np.zeros_like(cur), a, prev, (a + prev) >> 1, formats._paeth(a, prev, c)][ft]
        rows.append(ft)
        rows += bytes(((cur - pred) & 255).astype(np.uint8))
        prev = cur

    def chunk(t, b):
        return struct.pack(">I", len(b)) + t + b + struct.pack(">I", zlib.crc32(t + b) & 0xFFFFFFFF)

    data = zlib.compress(bytes(rows))
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 16, 0, 0, 0, 0)))
        f.write(chunk(b"IDAT", data[:len(data) // 2]) + chunk(b"IDAT", data[len(data) // 2:]) + chunk(b"IEND", b""))


def test_disparity_png_all_filters(tmp_path):
    rng = np.random.default_rng(1)
    img = rng.integers(0, 65536, size=(13, 9)).astype(np.uint16)
    p = str(tmp_path / "d.png")
    _png_with_filters(p, img)
    assert np.array_equal(formats.read_png_gray(p), img)
    d = formats.load_disparity(p)
    assert d.dtype == np.float32 and np.array_equal(d, img.astype(np.float32) / 256.0)
    formats.write_png_gray16(p, img)
    assert np.array_equal(formats.read_png_gray(p), img)


def test_disparity_from_flo_is_negated_first_channel(tmp_path):
    flow = np.stack([-np.arange(12, dtype=np.float32).reshape(3, 4), np.zeros((3, 4), np.float32)], -1)
    p = str(tmp_path / "d.flo")
    formats.save_flow(p, flow)
    assert np.array_equal(formats.load_disparity(p), np.arange(12, dtype=np.float32).reshape(3, 4))
    with pytest.raises(ValueError):
        formats.load_disparity(str(tmp_path / "d.exr"))


def test_pose_files(tmp_path):
    win = synth.make_window(32, 24, 4, seed=3)
    poses = np.zeros((4, 6))
    for i in range(4):
        R = win["Rs"][i].astype(np.float64)
        th = np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1))
        ax = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / (2 * np.sin(th))
        poses[i, :3], poses[i, 3:] = ax * th, win["ts"][i]
    T = formats.accumulate_poses(poses)
    assert len(T) == 5 and np.allclose(T[0], np.eye(4))
    # frame k camera-to-world: a point at the origin of camera k, mapped to world, then chained forward, is 0 again
    Xw = T[2][:3, 3]
    X = Xw.copy()
    for i in range(2):
        X = win["Rs"][i].astype(np.float64) @ X + win["ts"][i]
    assert np.abs(X).max() < 1e-5
    pk = str(tmp_path / "kitti.txt")
    formats.save_poses(pk, T, "KITTI")
    assert np.allclose(formats.load_poses_kitti(pk), np.stack(T))
    pt = str(tmp_path / "tartan.txt")
    formats.save_poses(pt, T, "TartanAir")
    rows = np.loadtxt(pt)
    assert rows.shape == (5, 7)
    for k in range(5):
        tz, tx, ty, qz, qx, qy, qw = rows[k]
        assert np.allclose([tx, ty, tz], T[k][:3, 3])
        x, y, z, w = qx, qy, qz, qw
        Rq = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                       [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                       [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert np.allclose(Rq, T[k][:3, :3], atol=1e-9)
    with pytest.raises(ValueError):
        formats.save_poses(pt, T, "EuRoC")
