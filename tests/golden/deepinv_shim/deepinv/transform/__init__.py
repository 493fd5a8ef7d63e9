import torch


class Shift(torch.nn.Module):
    """deepinv v0.2.0 transform/shift.py (n_trans=1): one random circular roll in H and W."""

    def __init__(self, n_trans=1, shift_max=1.0):
        super().__init__()
        self.n_trans = n_trans
        self.shift_max = shift_max

    def forward(self, x):
        H, W = x.shape[-2:]
        assert self.n_trans <= H - 1 and self.n_trans <= W - 1
        H_max, W_max = int(self.shift_max * H), int(self.shift_max * W)
        x_shift = torch.arange(-H_max, H_max)[torch.randperm(2 * H_max)][: self.n_trans]
        y_shift = torch.arange(-W_max, W_max)[torch.randperm(2 * W_max)][: self.n_trans]
        out = torch.cat([torch.roll(x, [sx, sy], [-2, -1]) for sx, sy in zip(x_shift, y_shift)], dim=0)
        return out


class Rotate(torch.nn.Module):
    """deepinv v0.2.0 transform/rotate.py (n_trans=1): one random integer-degree rotation."""

    def __init__(self, n_trans=1, degrees=360):
        super().__init__()
        self.n_trans, self.group_size = n_trans, degrees

    def forward(self, x):
        from torchvision.transforms.functional import rotate
        theta = torch.arange(0, 360)[1:][torch.randperm(359)]
        theta = theta[: self.n_trans]
        return torch.cat([rotate(x, float(_theta)) for _theta in theta])
