import torch, torch.nn.functional as F
torch.manual_seed(0)
for shape in ((2,32,8,8),(2,32,32,32),(2,32,64,64)):
    x = torch.randn(*shape, device='cuda')
    res = {}
    for fmt in ('nchw','nhwc'):
        xi = x.clone()
        if fmt == 'nhwc': xi = xi.contiguous(memory_format=torch.channels_last)
        xi.requires_grad_(True)
        for cip in (False, True):
            o = F.avg_pool2d(xi, 3, stride=2, padding=1, count_include_pad=cip)
            g = torch.arange(o.numel(), device='cuda', dtype=torch.float32).view_as(o).sin()
            (gx,) = torch.autograd.grad(o, xi, g)
            res[(fmt,cip)] = (o.detach(), gx)
    xc = x.cpu().requires_grad_(True)
    for cip in (False, True):
        o = F.avg_pool2d(xc, 3, stride=2, padding=1, count_include_pad=cip)
        g = torch.arange(o.numel(), dtype=torch.float32).view_as(o).sin()
        (gx,) = torch.autograd.grad(o, xc, g)
        for fmt in ('nchw','nhwc'):
            a = res[(fmt,cip)]
            print(shape, fmt, 'count_include_pad', cip, 'fwd err', (a[0].cpu()-o).abs().max().item(), 'bwd err', (a[1].cpu()-gx).abs().max().item())
