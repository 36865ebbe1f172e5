#!/usr/bin/env python
"""Regenerate tests/golden/ from the reference tree (run in the dev container only).

* deflate64_fixtures.bin/.json -- the raw deflate64 streams of the reference's test/data
  directory (inputs only; the reference ships no expected outputs), packed into one blob, with
  the output length / crc32 / adler32 obtained by decoding them with the oracle restatement
  (oracle/inflate.c) and cross-checked against SURVEY.md section 4.3.
* kat.json -- the known-answer vectors of the reference's own tests (file:line cited per entry).

/root/reference does not exist on the GPU box, hence the committed copies.
"""
import json
import os
import sys
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

REF = "/root/reference/test/data"


def main():
    blob = bytearray()
    entries = []
    seen = {}
    for name in sorted(os.listdir(REF)):
        if not name.endswith(".deflate64"):
            continue
        data = open(os.path.join(REF, name), "rb").read()
        if data in seen:
            off = seen[data]
        else:
            off = len(blob)
            seen[data] = off
            blob += data
        ret, out, used, _ = O.inflate(data, -16)
        assert ret == O.Z_STREAM_END and used == len(data), name
        entries.append({"name": name, "offset": off, "length": len(data), "out_len": len(out),
                        "crc32": zlib.crc32(out), "adler32": zlib.adler32(out)})
    open(os.path.join(HERE, "deflate64_fixtures.bin"), "wb").write(bytes(blob))
    json.dump({"source": "zlib-streams-ts test/data/*.deflate64", "fixtures": entries},
              open(os.path.join(HERE, "deflate64_fixtures.json"), "w"), indent=1)

    kat = {
        "crc32": [
            {"data_hex": b"hello".hex(), "init": 0, "expect": 0x3610A686,
             "ref": "test/coverage-targets/coverage-crc32.spec.ts:9-14"},
        ],
        "adler32": [
            {"data_hex": "05", "init": 0, "expect": 0x00050005,
             "ref": "test/coverage-targets/coverage-adler32.spec.ts:10-14"},
            {"data_hex": "010203", "init": 0, "expect": (10 << 16) | 6,
             "ref": "test/coverage-targets/coverage-adler32.spec.ts:16-21"},
        ],
        "inflate": [
            {"in_hex": "4b1c0500", "window_bits": -15, "ret": 1, "out_byte": 0x61, "out_len": 259,
             "ref": "test/inflate/test-inflate9-length-code-285.spec.ts:12-15,46-50"},
            {"in_hex": "4b1cfdff07a3e5030000", "window_bits": -16, "ret": 1, "out_byte": 0x61, "out_len": 66539,
             "ref": "test/inflate/test-inflate9-length-code-285.spec.ts:6-10,42-44"},
            {"in_hex": "000300fcff414243", "window_bits": -16, "ret": -5, "out_hex": b"ABC".hex(),
             "ref": "test/inflate/test-inflate9-stored-block.spec.ts:10-45"},
            {"in_hex": "010000ffff", "window_bits": 15, "ret_le": 0,
             "ref": "test/inflate/test-inftrees-infcover.spec.ts:11-39"},
            {"in_hex": "0400feff", "window_bits": -16, "ret_not": [0, 1],
             "ref": "test/inflate/test-inflate9-invalid-lengths.spec.ts:10-40"},
            {"in_hex": "", "window_bits": -16, "ret": -5,
             "ref": "test/inflate/test-inflate9-needmore.spec.ts:5-28"},
            {"in_hex": "0300", "window_bits": -15, "ret": 1, "out_hex": "",
             "ref": "test/round-trip/test-streams-empty-input.ts:32-35 (deflate-raw of empty input is 03 00)"},
        ],
        "headers": {
            "zlib": {"1": "7801", "6": "789c", "9": "78da", "ref": "src/mod/deflate/deflate.ts:754-770"},
            "gzip_os_byte": 255, "gzip_ref": "src/mod/deflate/deflate.ts:788-800",
        },
    }
    json.dump(kat, open(os.path.join(HERE, "kat.json"), "w"), indent=1)
    print("wrote", len(entries), "fixtures,", len(blob), "bytes")


if __name__ == "__main__":
    main()
