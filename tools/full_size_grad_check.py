import os, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/boosting-nerv_b200")
import bench
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
for cfg in sys.argv[1:]:
    model, args = bench.build_model(cfg)
    model = model.cuda().train()
    is_h = args.model == "HNeRV_Boost"
    fh, fw = [int(v) for v in args.fc_hw.split("_")]
    emb = torch.rand(1, 16, fh, fw, device="cuda", requires_grad=True) if is_h else None
    t = torch.tensor([0.37], dtype=torch.float64, device="cuda")
    res = {}
    target = None
    for mode in ("torch", "b200"):
        model.train_backend = mode
        model.zero_grad(set_to_none=True)
        if emb is not None: emb.grad = None
        img = (model.forward_decoder(emb, t) if is_h else model(t))[0]
        if target is None: target = torch.rand_like(img)
        loss = ((img - target) ** 2).mean() + 0.3 * (img - target).abs().mean()
        loss.backward()
        res[mode] = ({n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}, loss.item(), None if emb is None else emb.grad.clone())
        del img, loss
        torch.cuda.empty_cache()
    gt, lt, et = res["torch"]; gn, ln, en = res["b200"]
    worst = []
    for n in gt:
        a, b = gn[n].double().flatten(), gt[n].double().flatten()
        worst.append((((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item(), torch.nn.functional.cosine_similarity(a, b, dim=0).item(), n))
    worst.sort(reverse=True)
    print(cfg, "loss", lt, ln, "params", len(gt), len(gn))
    for w in worst[:5]: print("   ", w)
    print("   median rel", sorted(w[0] for w in worst)[len(worst)//2], "min cos", min(w[1] for w in worst))
    if et is not None: print("   emb grad rel", ((et-en).abs().max()/et.abs().max()).item())
