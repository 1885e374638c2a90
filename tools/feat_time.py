import os, sys, torch
sys.path.insert(0, "freesound-classification_b200")
from fsb200.runtime import FeatureExtractor
from ops.utils import make_mel_filterbanks
x = torch.randn(64, 441000, device="cuda") * 0.1
for desc, mode in (("mel_2048_1024_128", 2), ("stft_256_128", 1)):
    k = desc.split("_")
    fx = FeatureExtractor(int(k[1]), int(k[2]), make_mel_filterbanks(desc) if k[0] == "mel" else None)
    for _ in range(3): fx(x, mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fx(x, mode)
    e1.record(); torch.cuda.synchronize()
    print(desc, "ring", os.environ.get("FSB200_FEAT_RING", "1"), "%.1f us" % (e0.elapsed_time(e1) / 20 * 1e3))
