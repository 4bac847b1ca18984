#!/usr/bin/env python
"""Golden vectors for oracle/xxh3_oracle.c: XXH3_64bits digests from the xxhash Python module (a
binding of libxxhash 0.8.x; the reference pins xxHash 0.8.3 in subprojects/xxhash.wrap, and XXH3's
output is frozen since 0.8.0).  Inputs are the byte pattern `pattern(n)` below, so only (n, digest)
is stored.  Also two cacheKey vectors assembled exactly as Potential.hpp:324-336 does.

    python tests/golden/make_xxh3_golden.py > tests/golden/xxh3_vectors.json
"""
import json
import struct

import numpy as np
import xxhash


def pattern(n: int) -> bytes:
    return bytes(((i * 131 + n * 31 + (i >> 8) * 7) & 0xFF) for i in range(n))


def cache_key(pos, Z, box, pot_type, params_key):
    h = 0
    for blob in (pos.astype("<f8").tobytes(), Z.astype("<i4").tobytes(), box.astype("<f8").tobytes(),
                 struct.pack("<Q", pot_type), struct.pack("<Q", params_key)):
        h ^= xxhash.xxh3_64_intdigest(blob)
    return h


def main():
    lengths = list(range(0, 260)) + [511, 512, 513, 1023, 1024, 1025, 1087, 1088, 2048, 5232, 5236, 24000, 100003]
    out = {"source": f"xxhash module {xxhash.VERSION} / libxxhash {xxhash.XXHASH_VERSION}",
           "digests": [[n, str(xxhash.xxh3_64_intdigest(pattern(n)))] for n in lengths], "cache_keys": []}
    for n, seed in ((4, 1), (218, 2)):
        rng = np.random.default_rng(seed)
        pos = np.round(rng.random((n, 3)) * 20.0, 6)
        Z = np.where(np.arange(n) % 7 == 0, 1, 29).astype(np.int32)
        box = np.diag([15.3456, 21.702, 100.0])
        out["cache_keys"].append({"positions": pos.ravel().tolist(), "Z": Z.tolist(), "box": box.ravel().tolist(),
                                  "pot_type": 1, "params_key": 0x1234ABCD5678EF01,
                                  "key": str(cache_key(pos, Z, box, 1, 0x1234ABCD5678EF01))})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
