"""Worker of tests/test_multirank_gloo.py: rank r adapts its shard of the tasks of a golden meta-batch;
the meta-gradient is all-reduced (gloo on CPU here, NCCL on GPUs) and must equal the single-process one."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE); sys.path.insert(0, os.path.join(HERE, "emu"))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    device = sys.argv[1]
    if device == "cpu":
        import inject
        inject.enable()
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        torch.cuda.set_device(rank)
        device = f"cuda:{rank}"
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(device))
    import kernel_cases as kc
    from helpers import load_golden, params_of, relerr
    from fumi_b200.optim import FusedAdam
    name = "fumi_train_n5k5_d512_tanh"            # B = 2 tasks -> one per rank
    g, bank = load_golden(name)
    B = g["sup_ids"].shape[0]
    assert B % world == 0
    sl = slice(rank * B // world, (rank + 1) * B // world)
    gs = dict(g)
    for k in ("sup_ids", "qry_ids", "sup_y", "qry_y"):
        gs[k] = g[k][sl]
    model = kc.make_fumi(g, bank, params_of(g), device)
    opt = FusedAdam(model.parameters(), lr=float(g["lr"]), weight_decay=float(g["wd"]))
    eng = model._get_engine(device)
    res = eng.fumi_batch(model, kc._torchmeta_batch(gs, bank), steps=int(g["steps"]), step_size=float(g["alpha"]),
                         train=True)
    la = res["loss_acc"].cpu().numpy()
    assert abs(la[0] - float(g["loss"])) < 1e-4 * abs(float(g["loss"])), (la, float(g["loss"]))
    assert abs(la[1] - float(g["acc"])) < 1e-6
    assert np.array_equal(res["preds"].cpu().numpy(), g["preds"][sl])
    gmax = max(np.abs(g["grad:" + k]).max() for k in params_of(g))
    for k, p in model.named_parameters():
        got, ref = p.grad.cpu().numpy(), g["grad:" + k]
        if k == "hyper_net.2.bias":
            assert np.abs(got - ref).max() < 1e-4 * gmax
        else:
            assert relerr(got, ref) < 2e-4, (k, relerr(got, ref))
    opt.step()
    # every rank must hold identical parameters after the step
    flat = torch.cat([p.detach().reshape(-1).cpu() for p in model.parameters()])
    other = [torch.empty_like(flat) for _ in range(world)]
    if device == "cpu":
        dist.all_gather(other, flat)
        assert all(torch.equal(o, flat) for o in other)
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
