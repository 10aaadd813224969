"""Extract E. Heitz's 256-spp/64-dimension blue-noise sampler tables (sobol sequence, scrambling tile,
ranking tile; "A Low-Discrepancy Sampler that Distributes Monte Carlo Errors as a Blue Noise in Screen Space",
Heitz et al. 2019, https://eheitzresearch.wordpress.com/762-2) from the copy the reference ships in
lib/RenderSystem/common_bluenoise.h into a raw byte blob: sob[65536] | scr[131072] | rnk[131072].
The reference expands each byte to a uint at start-up (lib/rendercore_optix7/rendercore.cpp:247-254); so do we.
Run once in the build container; the blob is committed as data (it is third-party numeric data needed for
identical random numbers, not reference source)."""
import re, sys, hashlib
import numpy as np

src = open(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/lib/RenderSystem/common_bluenoise.h").read()
out = b""
for name, count in (("sob256_64", 8192), ("scr256_64", 16384), ("rnk256_64", 16384)):
    m = re.search(name + r"\[\d+\]\s*=\s*\{(.*?)\};", src, re.S)
    vals = [int(v, 16) for v in re.findall(r"0x[0-9a-fA-F]+", m.group(1))]
    assert len(vals) == count, (name, len(vals))
    out += np.array(vals, dtype="<u8").tobytes()
assert len(out) == 65536 + 2 * 131072
open("lighthouse2_b200/data/heitz_bluenoise_256spp.bin", "wb").write(out)
print(len(out), hashlib.md5(out).hexdigest())
