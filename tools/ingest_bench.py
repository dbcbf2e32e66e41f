"""Frame-ingest throughput: the reference's reader (PIL open -> RGB -> resize 224, one host thread) vs
the GPU path (vsb_jpeg_decode_resize: Huffman on W host threads, pixels on the device) on 640x360 4:2:0 JPEGs."""
import io, json, os, sys, time
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from PIL import Image
from common import jpeg_bytes
from vidsitu_b200.jpeg import JpegDecoder

K, N = 48, int(os.environ.get("FRAMES", "960"))
datas = [jpeg_bytes(360, 640, 95, 2, "noisy" if i % 3 else "smooth", seed=i) for i in range(K)]
res = {"frame": "640x360 YCbCr 4:2:0 quality 95", "mean_jpeg_kb": round(sum(map(len, datas)) / K / 1024, 1), "frames": N}
t0 = time.time()
for i in range(N // 4):
    np.array(Image.open(io.BytesIO(datas[i % K])).convert("RGB").resize((224, 224)))
res["pil_1_thread_fps"] = round((N // 4) / (time.time() - t0), 1)
out = torch.empty((N, 224, 224, 3), dtype=torch.uint8, device="cuda")
for w in (1, 2, 4, 8, 16):
    decs = [(JpegDecoder(1920, 1088), torch.cuda.Stream()) for _ in range(w)]
    def work(t):
        dec, st = decs[t]
        with torch.cuda.stream(st):
            for i in range(t, N, w):
                dec.decode_resize(datas[i % K], out[i])
        st.synchronize()
    with ThreadPoolExecutor(w) as pool:
        list(pool.map(work, range(w)))      # warm-up
        torch.cuda.synchronize()
        t0 = time.time()
        list(pool.map(work, range(w)))
        torch.cuda.synchronize()
    res[f"gpu_path_{w}_threads_fps"] = round(N / (time.time() - t0), 1)
    del decs
from vidsitu_b200.jpeg import JpegBatchDecoder
bd = JpegBatchDecoder()
for nb in (160, 480, 960, 1920):
    batch = [datas[i % K] for i in range(nb)]
    o = torch.empty((nb, 224, 224, 3), dtype=torch.uint8, device="cuda")
    outs = list(o)
    assert all(bd.decode_resize(batch, outs))     # warm-up (workspaces grow)
    torch.cuda.synchronize()
    t0 = time.time()
    reps = max(1, 1920 // nb)
    for _ in range(reps):
        bd.decode_resize(batch, outs)
    torch.cuda.synchronize()
    res[f"gpu_huffman_batch_{nb}_fps"] = round(nb * reps / (time.time() - t0), 1)
    res[f"gpu_huffman_batch_{nb}_bit_exact"] = bool(np.array_equal(o[5].cpu().numpy(), np.array(Image.open(io.BytesIO(batch[5])).convert("RGB").resize((224, 224)))))
ref = np.array(Image.open(io.BytesIO(datas[5])).convert("RGB").resize((224, 224)))
res["bit_exact"] = bool(np.array_equal(out[5].cpu().numpy(), ref))
res["host_cpus"] = os.cpu_count()
print(json.dumps(res))
