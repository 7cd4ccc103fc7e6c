"""BASELINE config 5 (S3DIS-shaped full-scene inference): Res16UNet34C, 5 cm voxels, eval-mode BatchNorm, 13 classes,
forward only, one synthetic room per call (`downstream/semseg/lib/test.py:95-117`).  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pointcontrast_b200 import me, synth  # noqa: E402
from pointcontrast_b200.config import default_config  # noqa: E402
from pointcontrast_b200.model import load_model  # noqa: E402


def main():
    cfg = default_config(["net.normalize_feature=False"])
    net = load_model("Res16UNet34C")(3, 13, cfg, D=3).cuda().eval()
    scenes = [synth.synth_scene(s) for s in range(3)]
    dev = [(torch.from_numpy(s["feats"]).cuda(), torch.from_numpy(s["coords"]).cuda()) for s in scenes]
    host = [(torch.from_numpy(s["feats"]).pin_memory(), torch.from_numpy(s["coords"]).pin_memory()) for s in scenes]

    def run(batches, n, to_host):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for i in range(n):
            f, c = batches[i % len(batches)]
            with torch.no_grad():
                out = net(me.SparseTensor(f, coords=c).to("cuda")).F
                pred = out.argmax(1)
            if to_host:
                pred = pred.cpu()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    run(dev, 3, False)
    ms = run(dev, 10, False)
    ms_e2e = run(host, 10, True)
    nvox = sum(len(s["coords"]) for s in scenes) / len(scenes)
    print(json.dumps({"config": "S3DIS-shape inference, Res16UNet34C, 5cm, eval BN, 13 classes, 1 scene/call", "voxels_per_scene": nvox,
                      "ms_per_scene": ms, "scenes_per_s": 1e3 / ms, "voxels_per_s": nvox * 1e3 / ms,
                      "e2e_ms_per_scene": ms_e2e, "e2e_scenes_per_s": 1e3 / ms_e2e}))


if __name__ == "__main__":
    main()
