"""Recompress the stored-deflate PNGs written by include/rtx/image_io.h (no compression) with
zlib level 9, for committing gallery images. Usage: python tools/recompress_png.py in.png out.png"""
import struct
import sys
import zlib


def chunks(raw):
    pos = 8
    while pos < len(raw):
        n, = struct.unpack(">I", raw[pos:pos + 4])
        yield raw[pos + 4:pos + 8], raw[pos + 8:pos + 8 + n]
        pos += 12 + n


def chunk(kind, data):
    return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data) & 0xffffffff)


raw = open(sys.argv[1], "rb").read()
assert raw[:8] == b"\x89PNG\r\n\x1a\n"
ihdr, idat = None, b""
for kind, data in chunks(raw):
    if kind == b"IHDR":
        ihdr = data
    elif kind == b"IDAT":
        idat += data
out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", zlib.compress(zlib.decompress(idat), 9)) + chunk(b"IEND", b"")
open(sys.argv[2], "wb").write(out)
print(sys.argv[2], len(raw), "->", len(out))
