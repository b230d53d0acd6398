"""Golden vectors for the image pre-/post-processing row: outputs of the REAL libraries the reference calls --
PIL Image.resize(BICUBIC) (data/base_dataset.py:150-172), torchvision ToTensor + Normalize (:143-148) and the reference's
own util.tensor2im (util/util.py:98-131, imported from /root/reference when it is mounted) -- on seeded inputs.
    python oracle/gen_golden_imageio.py        # writes tests/golden/imageio.npz"""
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

rng = np.random.RandomState(2024)
out = {}
cases = [((37, 53), (20, 14)), ((64, 48), (128, 96)), ((75, 50), (128, 192)), ((97, 131), (97, 64)), ((50, 50), (16, 50)),
         ((256, 171), (192, 128)), ((9, 7), (3, 2)), ((31, 200), (512, 80))]
names = []
for i, ((h, w), (ow, oh)) in enumerate(cases):
    # smooth + noise so that both ringing (clipping) and flat regions occur
    base = rng.rand(max(h // 8, 1), max(w // 8, 1), 3)
    img = np.asarray(Image.fromarray((base * 255).astype(np.uint8)).resize((w, h), Image.BILINEAR)).astype(np.int32)
    img = np.clip(img + rng.randint(-60, 60, size=img.shape), 0, 255).astype(np.uint8)
    res = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
    out["resize/%d/in" % i], out["resize/%d/size" % i], out["resize/%d/out" % i] = img, np.array([ow, oh]), res
    names.append(str(i))
out["resize/names"] = np.array(names)

import torchvision.transforms as T  # noqa: E402
img = out["resize/0/in"]
tt = T.Compose([T.ToTensor(), T.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])(Image.fromarray(img)).numpy()
out["totensor/in"], out["totensor/out"] = img, tt
allv = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2)
out["totensor/all_in"] = allv
out["totensor/all_out"] = T.Compose([T.ToTensor(), T.Normalize((0.5,) * 3, (0.5,) * 3)])(Image.fromarray(allv)).numpy()

x = (rng.rand(3, 40, 56).astype(np.float32) * 2.6 - 1.3)
x[0, 0, :8] = [-1.0, 1.0, 0.0, 0.999999, -0.999999, 1.0000001, np.float32(2 / 255 * 100 - 1), np.float32(1 / 3)]
try:
    sys.path.insert(0, "/root/reference")
    from oracle import ref_harness  # noqa: F401  (stubs for dominate / visdom etc.)
    ref_harness.install()
    import util as ref_util
    im = ref_util.tensor2im(__import__("torch").from_numpy(x), tile=False)
    src = "reference util.tensor2im"
except Exception as e:  # noqa: BLE001
    im = np.clip((np.transpose(x, (1, 2, 0)) + 1) / 2.0 * 255.0, 0, 255).astype(np.uint8)
    src = "numpy expression of util/util.py:122-128 (reference import failed: %r)" % (e,)
out["tensor2im/in"], out["tensor2im/out"], out["tensor2im/source"] = x, im, np.array(src)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "imageio.npz"), **out)
print("wrote imageio.npz;", src, {k: v.shape for k, v in out.items() if k.endswith("/out")})
