"""Tensor-parallel decode (BASELINE config c5; spatialrgpt_b200/tensor_parallel.py).  The sharded decode step is exercised end to end
on ONE GPU with world = 1 (every shard is the whole matrix, the collectives are no-ops: the kernel sequence, the fp32 partial-sum /
residual path, the vocabulary-parallel arg max and the CUDA graph must reproduce the plain decoder's ids exactly), per-rank slices
are emulated in tests/test_gpu_ops.py::test_tp_shards_reproduce_the_full_decode_ops, the host logic runs over gloo in
tests/test_dist_cpu.py, and with >= 2 GPUs the real thing runs under torchrun (tools/tp_run.py --check: TP-2 ids == TP-1 ids)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from oracle import srgpt_oracle as O
from tests.golden.make_golden import CASES
from tests.test_gpu_pipeline import build_model

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_tp_world1_equals_plain_decoder():
    from spatialrgpt_b200.tensor_parallel import TPLlamaDecoder

    kw, n_regions, t_text, kind, n_new, _ = CASES["tiny_masks_gqa"]
    oc, sd, model = build_model(kw, 5)
    ids, im, de, mk = O.synth_request(oc, n_regions, t_text, seed=1234, kind=kind)
    a = dict(images=im.to(DEV), depths=de.to(DEV), masks=[m.to(DEV) for m in mk], do_sample=False, max_new_tokens=20)
    ref = model.generate(ids.to(DEV), **a)[0].tolist()
    model.llm = TPLlamaDecoder(model.config.llama, model.weights.llama, 0, 1, max_seq_len=512)
    assert model.generate(ids.to(DEV), **a)[0].tolist() == ref
    assert model.generate(ids.to(DEV), use_cuda_graph=False, **a)[0].tolist() == ref
    assert model.generate(ids.to(DEV), eos_token_id=ref[4], **a)[0].tolist() == ref[: ref.index(ref[4]) + 1]
    with pytest.raises(NotImplementedError):
        model.generate(ids.to(DEV), images=a["images"], depths=a["depths"], masks=a["masks"], do_sample=True, temperature=0.8, max_new_tokens=4)


@pytest.mark.skipif(torch.cuda.device_count() < 2 or os.environ.get("SRGPT_RUN_TP2_TEST") != "1",
                    reason="needs 2 GPUs and SRGPT_RUN_TP2_TEST=1 (a torchrun child of 2 ranks; run by hand: gpurun --gpus 2, tools/gpu_job_tp.sh 2)")
def test_tp2_ids_equal_tp1_under_torchrun():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(root, "tools", "tp_run.py"), "--check"], capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0 and '"ok": true' in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
