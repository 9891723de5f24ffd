import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
pkg = importlib.import_module("prostate-cancer-multimodal-segmentation_b200")
import unet3d_oracle as oracle
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
for f, shape, ncls in ((48, (2, 5, 32, 32, 32), 1), (16, (1, 5, 16, 48, 24), 3), (80, (1, 5, 16, 16, 32), 1)):
    torch.manual_seed(0)
    model = pkg.UNet3D(5, ncls, init_features=f)
    sd = {k: v.detach().clone().to(dev) for k, v in model.state_dict().items()}
    model = model.to(dev).train()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(*shape, generator=g).to(dev)
    y = (torch.rand(shape[0], ncls, *shape[2:], generator=g) < 0.1).float().to(dev)
    logits = model(x)
    loss = pkg.BCEDiceLoss()(logits, y)
    loss.backward()
    torch.cuda.synchronize()
    work = {k: v.clone() for k, v in sd.items()}
    names = oracle.param_names(sd)
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    work.update(leaves)
    ol = oracle.unet3d_forward(x, work, training=True)
    oloss = oracle.bce_dice_loss(ol.float(), y)
    grads = dict(zip(names, torch.autograd.grad(oloss, [leaves[k] for k in names])))
    rel = ((logits - ol).norm() / ol.norm()).item()
    gname = "up4.conv.conv.3.weight"
    grel = ((dict(model.named_parameters())[gname].grad - grads[gname]).norm() / grads[gname].norm()).item()
    g0 = "inc.conv.0.weight"
    g0rel = ((dict(model.named_parameters())[g0].grad - grads[g0]).norm() / grads[g0].norm()).item()
    print(f"base {f} shape {shape} classes {ncls}: loss {loss.item():.5f} vs {oloss.item():.5f}, logits rel-L2 {rel:.2e}, "
          f"{gname} grad rel-L2 {grel:.2e}, {g0} grad rel-L2 {g0rel:.2e}")
