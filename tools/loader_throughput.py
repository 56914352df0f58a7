"""Image-loader throughput on JPEG files: full-size photos -> 1024-px LANCZOS thumbnail -> K1 CLAHE
transform, without a backbone, for the three arrangements a user can pick:
  host      DataLoader workers decode (PIL) AND thumbnail (Pillow) -- the reference's arrangement; 2.4 MB per image uploaded
  k5        workers only decode (PIL); crop / thumbnail on the device (K5, bit-identical to Pillow); 21 MB per image uploaded
  nvjpeg    JPEG bit streams to the GPU (gdt_jpeg_decode_batch: nvJPEG library, hardware JPEG engines when available; not
            bit-identical to libjpeg) + K5, batches of 8 / 64 files
Prints one JSON line. JPEGs are synthetic 3072x2304 photos (smooth sinusoid + noise family, quality 90) written to a
temporary directory first.
    python tools/loader_throughput.py [n_images] [workers]"""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from PIL import Image

from gandtr_b200.extract import _Decode
from gandtr_b200.loader import DeviceImageLoader
from tests.util import synth_image

GH, GW, IMSIZE = 2304, 3072, 1024


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    workers = int(sys.argv[2]) if len(sys.argv) > 2 else min(16, os.cpu_count() or 4)
    dev = torch.device("cuda", 0)
    tmp = tempfile.mkdtemp(prefix="gdt_jpeg_")
    base = [synth_image(500 + i, GH, GW, "smooth") for i in range(4)]
    paths = []
    for i in range(n):
        p = os.path.join(tmp, "photo_%03d.jpg" % i)
        Image.fromarray(np.roll(base[i % 4], 37 * i, axis=1)).save(p, quality=90)
        paths.append(p)
    jpeg_mb = sum(os.path.getsize(p) for p in paths) / n / 1e6
    from gandtr_b200 import hub
    net = hub.gem_vgg16_hedngan(pretrained=False)
    transform = net.transform
    out = {"images": n, "photo": "%dx%d JPEG q90, %.2f MB on disk" % (GW, GH, jpeg_mb), "workers": workers, "thumbnail": IMSIZE}

    def run(mode):
        geometry = DeviceImageLoader(imsize=IMSIZE, device=dev, decode="nvjpeg" if mode.startswith("nvjpeg") else "pil")
        t0 = time.time()
        done, pending = 0, []

        def flush():
            nonlocal pending, done
            if pending:
                transform.batch(torch.stack(pending))
                done += len(pending)
                pending = []
        if mode.startswith("nvjpeg"):
            bs = 64 if mode.endswith("64") else 8
            for a in range(0, n, bs):                   # batched GPU decode + one K5 launch pair per batch of files
                pending = geometry.load_batch(paths[a:a + bs])
                flush()
        else:
            ds = _Decode(paths, IMSIZE, None, device_resize=(mode == "k5"))
            dl = torch.utils.data.DataLoader(ds, batch_size=None, shuffle=False, num_workers=workers)
            for item in dl:
                if mode == "k5":
                    img, _ = item
                    img = geometry.resize(img)
                else:
                    img = item.pin_memory().to(dev, non_blocking=True)
                pending.append(img)
                if len(pending) == 8:
                    flush()
        flush()
        torch.cuda.synchronize()
        return n / (time.time() - t0)

    for mode in ("host", "k5", "nvjpeg", "nvjpeg64"):
        try:
            run(mode)                                   # warm-up: worker start-up, plans, gdt_init
            out[mode + "_images_per_s"] = round(run(mode), 1)
        except Exception as e:                          # noqa: BLE001 -- e.g. a torchvision build without nvjpeg
            out[mode + "_images_per_s"] = "unavailable: %s" % str(e)[:120]
    from gandtr_b200 import _lib
    import ctypes
    st = (ctypes.c_int * 4)()
    _lib.load().gdt_debug_jpeg_status(st)
    out["nvjpeg_backend_of_last_batch"] = {1: "hardware JPEG engines", 2: "GPU-hybrid", 3: "default (host Huffman threads)"}.get(_lib.load().gdt_debug_jpeg_last_backend(), "none")
    out["nvjpeg_status_hw_create_init_decode_fallback"] = list(st)
    print(json.dumps(out))
    for p in paths:
        os.remove(p)
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
