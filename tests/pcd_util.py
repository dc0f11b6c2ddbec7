"""Test-side PCD v0.7 writers (ascii / binary / binary_compressed) used to exercise the room input path.  The compressed writer
emits real LZF back references (greedy longest match in a short window), so the decoder's copy paths are covered."""
import numpy as np


def lzf_compress(data: bytes, window: int = 512) -> bytes:
    out = bytearray()
    lit = bytearray()
    i, n = 0, len(data)

    def flush():
        nonlocal lit
        while lit:
            run = lit[:32]
            out.append(len(run) - 1)
            out.extend(run)
            lit = lit[32:]

    while i < n:
        best_len, best_off = 0, 0
        if i + 3 <= n:
            lo = max(0, i - window)
            key = data[i:i + 3]
            j = data.rfind(key, lo, i + 2)
            tries = 0
            while j != -1 and j < i and tries < 8:
                ln = 3
                while i + ln < n and ln < 264 and data[j + ln] == data[i + ln]:
                    ln += 1
                if ln > best_len:
                    best_len, best_off = ln, i - j
                j = data.rfind(key, lo, j + 2) if j > lo else -1
                tries += 1
        if best_len >= 3:
            flush()
            ln, off = best_len - 2, best_off - 1
            if ln < 7:
                out.append((ln << 5) | (off >> 8))
            else:
                out.append((7 << 5) | (off >> 8))
                out.append(ln - 7)
            out.append(off & 0xFF)
            i += best_len
        else:
            lit.append(data[i])
            i += 1
    flush()
    return bytes(out)


def write_pcd(path, xyz, rgb=None, normals=None, kind="binary", rgb_type="F", extra_front=False):
    """xyz float32 [n,3]; rgb uint8 [n,3] packed as 0x00RRGGBB in a 4-byte F or U field; normals float32 [n,4] (nx ny nz curvature).
    extra_front puts a 1-byte `tag` field first, so every later field sits at an unaligned offset."""
    xyz = np.asarray(xyz, np.float32).reshape(-1, 3)
    n = len(xyz)
    cols, fields, sizes, types = [], [], [], []
    if extra_front:
        cols.append((np.arange(n) % 251).astype(np.uint8)); fields.append("tag"); sizes.append(1); types.append("U")
    for c, a in enumerate("xyz"):
        cols.append(xyz[:, c].copy()); fields.append(a); sizes.append(4); types.append("F")
    if rgb is not None:
        rgb = np.asarray(rgb, np.uint32).reshape(-1, 3)
        bits = ((rgb[:, 0] << 16) | (rgb[:, 1] << 8) | rgb[:, 2]).astype(np.uint32)
        cols.append(bits.view(np.float32) if rgb_type == "F" else bits); fields.append("rgb"); sizes.append(4); types.append(rgb_type)
    if normals is not None:
        normals = np.asarray(normals, np.float32).reshape(-1, 4)
        for c, a in enumerate(("normal_x", "normal_y", "normal_z", "curvature")):
            cols.append(normals[:, c].copy()); fields.append(a); sizes.append(4); types.append("F")
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\n"
           f"FIELDS {' '.join(fields)}\nSIZE {' '.join(map(str, sizes))}\nTYPE {' '.join(types)}\nCOUNT {' '.join('1' for _ in fields)}\n"
           f"WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA {kind}\n").encode()
    with open(path, "wb") as fh:
        fh.write(hdr)
        if kind == "ascii":
            for i in range(n):
                toks = []
                for col, t in zip(cols, types):
                    v = col[i]
                    toks.append(repr(float(v)) if t == "F" and col.dtype == np.float32 and np.isfinite(v) else str(int(v)) if t != "F" else repr(float(v)))
                fh.write((" ".join(toks) + "\n").encode())
        elif kind == "binary":
            dt = np.dtype([(f, c.dtype) for f, c in zip(fields, cols)])
            rec = np.empty(n, dt)
            for f, c in zip(fields, cols):
                rec[f] = c
            fh.write(rec.tobytes())
        else:
            soa = b"".join(c.tobytes() for c in cols)
            comp = lzf_compress(soa)
            fh.write(np.array([len(comp), len(soa)], np.uint32).tobytes())
            fh.write(comp)
