"""Deterministic synthetic multi-stream scene generator.

Produces, per tick, the padded detection batch the tracker consumes -- the state of the pipeline
right after NMS and the re-ID encoder (deepdish.py:1014): integer tlwh boxes, unique scores, labels
and 128-d appearance features for S independent camera streams.  Written with torch ops so the
same code generates small CPU batches for the parity tests and device-resident batches for
``bench.py`` (``device='cuda'``); a fixed seed gives a fixed scene on a given device type.

Objects move with constant velocity and reflect at the frame border (so paths cross the default
count-line x = W/2), are missed with probability ``miss_prob`` (exercises cascade levels and
deletion), are re-spawned with a new identity with probability ``respawn_prob`` (exercises track
ageing / id allocation) and every frame carries a few clutter boxes with random features
(exercises tentative-track deletion).
"""
import torch


class SceneBatch:
    """One tick of detections for S streams (padded to Dmax)."""
    __slots__ = ("tlwh", "conf", "label", "feat", "count")

    def __init__(self, tlwh, conf, label, feat, count):
        self.tlwh, self.conf, self.label, self.feat, self.count = tlwh, conf, label, feat, count

    def to(self, device, non_blocking=False):
        return SceneBatch(*(getattr(self, k).to(device, non_blocking=non_blocking)
                            for k in self.__slots__))

    def pin(self):
        return SceneBatch(*(getattr(self, k).pin_memory() for k in self.__slots__))

    def stream(self, s):
        """Host view of one stream's detections as plain numpy arrays (for the oracle)."""
        n = int(self.count[s])
        return (self.tlwh[s, :n].cpu().numpy(), self.conf[s, :n].cpu().numpy(),
                self.label[s, :n].cpu().numpy(), self.feat[s, :n].cpu().numpy())

    def nbytes(self):
        return sum(getattr(self, k).numel() * getattr(self, k).element_size() for k in self.__slots__)


class Scene:
    def __init__(self, n_streams, n_objects, dmax, n_labels=1, width=640, height=480, seed=0,
                 miss_prob=0.1, respawn_prob=0.004, clutter_mean=2.0, label_noise=0.05,
                 feat_noise=0.02, feat_dim=128, device="cpu"):
        self.S, self.N, self.Dmax = n_streams, n_objects, dmax
        self.C, self.W, self.H = n_labels, width, height
        self.miss_prob, self.respawn_prob = miss_prob, respawn_prob
        self.clutter_mean, self.label_noise, self.feat_noise = clutter_mean, label_noise, feat_noise
        self.F = feat_dim
        self.device = torch.device(device)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(seed)
        S, N = self.S, self.N
        self.pos = self._rand(S, N, 2) * torch.tensor([600.0, 400.0], device=self.device)
        self.vel = self._randn(S, N, 2) * 3.0
        self.size = torch.stack([20 + 20 * self._rand(S, N), 40 + 60 * self._rand(S, N)], dim=-1)
        self.ident = self._unit(self._randn(S, N, self.F))
        self.label = torch.randint(0, self.C, (S, N), generator=self.gen, device=self.device)
        self.max_clutter = max(0, dmax - n_objects)

    def _rand(self, *shape):
        return torch.rand(*shape, generator=self.gen, device=self.device, dtype=torch.float32)

    def _randn(self, *shape):
        return torch.randn(*shape, generator=self.gen, device=self.device, dtype=torch.float32)

    @staticmethod
    def _unit(x):
        return x / x.norm(dim=-1, keepdim=True)

    def step(self):
        S, N, D, dev = self.S, self.N, self.Dmax, self.device
        # motion with reflection
        self.pos = self.pos + self.vel
        lim = torch.tensor([600.0, 400.0], device=dev)
        over, under = self.pos > lim, self.pos < 0
        self.pos = torch.where(over, 2 * lim - self.pos, torch.where(under, -self.pos, self.pos))
        self.vel = torch.where(over | under, -self.vel, self.vel)
        # re-spawn a few objects with a new identity
        rs = self._rand(S, N) < self.respawn_prob
        if bool(rs.any()):
            npos = self._rand(S, N, 2) * lim
            nid = self._unit(self._randn(S, N, self.F))
            self.pos = torch.where(rs[..., None], npos, self.pos)
            self.vel = torch.where(rs[..., None], self._randn(S, N, 2) * 3.0, self.vel)
            self.ident = torch.where(rs[..., None], nid, self.ident)
        M = N + self.max_clutter
        # candidate table: N objects then clutter slots
        tl = self.pos + self._randn(S, N, 2)
        wh = self.size + self._randn(S, N, 2)
        feat = self._unit(self.ident + self.feat_noise * self._randn(S, N, self.F))
        lab = self.label
        flip = self._rand(S, N) < self.label_noise
        lab = torch.where(flip, torch.randint(0, self.C, (S, N), generator=self.gen, device=dev), lab)
        present = self._rand(S, N) >= self.miss_prob
        if self.max_clutter > 0:
            K = self.max_clutter
            ctl = self._rand(S, K, 2) * lim
            cwh = torch.stack([20 + 20 * self._rand(S, K), 40 + 60 * self._rand(S, K)], dim=-1)
            cfeat = self._unit(self._randn(S, K, self.F))
            clab = torch.randint(0, self.C, (S, K), generator=self.gen, device=dev)
            p_clutter = min(1.0, self.clutter_mean / K)
            cpres = self._rand(S, K) < p_clutter
            tl, wh = torch.cat([tl, ctl], 1), torch.cat([wh, cwh], 1)
            feat, lab = torch.cat([feat, cfeat], 1), torch.cat([lab, clab], 1)
            present = torch.cat([present, cpres], 1)
        # integer boxes inside the frame, w,h >= 1 (deepdish.py:950-951 would leave these unchanged)
        x = tl[..., 0].clamp(0, self.W - 2).floor()
        y = tl[..., 1].clamp(0, self.H - 2).floor()
        w = torch.minimum(wh[..., 0].floor().clamp(min=1), self.W - x)
        h = torch.minimum(wh[..., 1].floor().clamp(min=1), self.H - y)
        # unique scores: a random permutation picks distinct bins of [0.5, 1)
        perm = torch.argsort(self._rand(S, M), dim=1).to(torch.float32)
        score = 0.5 + (perm + 0.25 + 0.5 * self._rand(S, M)) / (2.0 * M)
        # present detections first, in descending score (the order NMS hands them on)
        key = torch.where(present, score, torch.full_like(score, -1.0))
        order = torch.argsort(key, dim=1, descending=True)[:, :D]
        count = present.sum(dim=1).clamp(max=D).to(torch.int32)

        def take(a):
            idx = order
            while idx.dim() < a.dim():
                idx = idx[..., None]
            return torch.gather(a, 1, idx.expand(-1, -1, *a.shape[2:]))

        tlwh = torch.stack([take(x), take(y), take(w), take(h)], dim=-1).to(torch.float64)
        valid = torch.arange(D, device=dev)[None, :] < count[:, None]
        tlwh = torch.where(valid[..., None], tlwh, torch.zeros_like(tlwh))
        return SceneBatch(tlwh.contiguous(), take(score).contiguous(),
                          take(lab).to(torch.int32).contiguous(), take(feat).contiguous(), count)
