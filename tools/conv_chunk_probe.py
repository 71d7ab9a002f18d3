"""GPU probe: does running the first convolutions of enc_b on sub-batches of notes keep the big
intermediate activations (4.2 MB per note after the first convolution) in the 126 MB L2 instead of
sending them through HBM?  Library convolutions only (cuDNN through torch), replayed from CUDA
graphs so that launch overhead does not decide the answer."""
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from interactive_spectrogram_inpainting_b200.vqvae import vqvae as vq  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
B = 444
model = vq.VQVAE(in_channel=2, resolution_factors={"bottom": 16, "top": 2}, adapt_quantized_durations=False)
model = model.to(dev).eval().to(memory_format=torch.channels_last)
enc = model.enc_b
x = torch.randn(B, 512, 64, 8, device=dev).permute(0, 3, 1, 2)      # space-to-depth blocks, channels_last
blocks = list(enc.blocks)
w1 = enc.space_to_depth_weight()


def head(xs, depth):
    """conv1 (3x3 over the blocks) + ReLU, then the next depth-1 strided convolutions + ReLU."""
    h = torch.cudnn_convolution_relu(xs, w1, blocks[0].bias, (1, 1), (1, 1), (1, 1), 1)
    for i in range(1, depth):
        c = blocks[2 * i]
        h = torch.cudnn_convolution_relu(h, c.weight, c.bias, c.stride, c.padding, c.dilation, c.groups)
    return h


def timed_graph(fn, iters=10):
    with torch.no_grad():
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = fn()
        g.replay()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(iters):
            g.replay()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / iters, out


for depth in (2, 4):
    ms_full, ref = timed_graph(lambda: head(x, depth))
    print(f"depth {depth}: whole batch {ms_full:.3f} ms, output {tuple(ref.shape)}", flush=True)
    for chunk in (4, 8, 12, 16, 24, 37, 74):
        out = torch.empty_like(ref)

        def chunked():
            for i in range(0, B, chunk):
                out[i:i + chunk].copy_(head(x[i:i + chunk], depth))
            return out
        ms, got = timed_graph(chunked)
        same = torch.equal(got, ref)
        err = (got - ref).abs().max().item()
        print(f"  chunks of {chunk:3d}: {ms:.3f} ms  ({ms_full - ms:+.3f})  identical={same} max|diff|={err:.2e}", flush=True)
