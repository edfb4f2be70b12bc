"""Minimal FITS binary-table reader (test infrastructure): enough to read the EVENTS extension the stock marx2fits
writes through jdfits, and the CALDB sub-pixel table it reads (fixed-width columns of types I, J, E, D, X, A with a
repeat count).  No astropy in this image."""
import re

import numpy as np

_TFORM = re.compile(r"^\s*(\d*)([LXBIJKAED])")
_NP = {"B": "u1", "I": ">i2", "J": ">i4", "K": ">i8", "E": ">f4", "D": ">f8", "L": "u1", "A": "S1"}


def _read_header(buf, pos):
    cards = {}
    while True:
        block = buf[pos:pos + 2880]
        if len(block) < 2880:
            return None, pos
        pos += 2880
        end = False
        for i in range(0, 2880, 80):
            card = block[i:i + 80].decode("latin-1")
            key = card[:8].strip()
            if key == "END":
                end = True
                break
            if card[8:10] != "= ":
                continue
            val = card[10:]
            if val.lstrip().startswith("'"):
                m = re.match(r"\s*'((?:[^']|'')*)'", val)
                v = m.group(1).rstrip() if m else val.strip()
            else:
                v = val.split("/")[0].strip()
                try:
                    v = int(v)
                except ValueError:
                    try:
                        v = float(v.replace("D", "E"))
                    except ValueError:
                        pass
            cards.setdefault(key, v)
        if end:
            return cards, pos


def read_bintable(path, extname):
    """-> (dict column name -> numpy array in native byte order, header dict) of the first extension named extname"""
    buf = open(path, "rb").read()
    pos = 0
    while True:
        hdr, pos = _read_header(buf, pos)
        if hdr is None:
            raise KeyError("%s: no extension %s" % (path, extname))
        naxis = hdr.get("NAXIS", 0)
        size = 0
        if naxis:
            size = abs(hdr["BITPIX"]) // 8
            for k in range(1, naxis + 1):
                size *= hdr["NAXIS%d" % k]
            size += hdr.get("PCOUNT", 0)
        data_pos, pos = pos, pos + (size + 2879) // 2880 * 2880
        if hdr.get("XTENSION", "") != "BINTABLE" or str(hdr.get("EXTNAME", "")).strip() != extname:
            continue
        fields = []
        for k in range(1, hdr["TFIELDS"] + 1):
            m = _TFORM.match(str(hdr["TFORM%d" % k]))
            rep = int(m.group(1)) if m.group(1) else 1
            t = m.group(2)
            name = str(hdr.get("TTYPE%d" % k, "COL%d" % k)).strip()
            if t == "X":
                fields.append((name, "u1", ((rep + 7) // 8,)))
            elif t == "A":
                fields.append((name, "S%d" % rep))
            elif rep == 1:
                fields.append((name, _NP[t]))
            else:
                fields.append((name, _NP[t], (rep,)))
        dt = np.dtype(fields)
        assert dt.itemsize == hdr["NAXIS1"], (dt.itemsize, hdr["NAXIS1"])
        rows = np.frombuffer(buf, dtype=dt, count=hdr["NAXIS2"], offset=data_pos)
        out = {}
        for name in dt.names:
            a = rows[name]
            out[name] = a.astype(a.dtype.newbyteorder("=")) if a.dtype.kind in "iuf" else a.copy()
        return out, hdr
