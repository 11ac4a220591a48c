"""The reference's train step driven through the reference's OWN ``Model`` (oracle/_ref, or
/root/reference in the build container).  TEST / BASELINE INFRASTRUCTURE ONLY.

``train.py`` of the reference cannot be imported on Python >= 3.7 (``async=True``,
code/dmcnet/train.py:226), so the loop body is restated here around the unmodified ``model.py``:
optimizer wiring code/dmcnet/train.py:121-142 (GAN :122-153), step body :221-266 (GAN :237-372).
The same restatement is what oracle/pin_against_reference.py holds the oracle to, bit for bit.
``bench.py`` times this class for ``cpu_baseline`` / ``--impl reference`` (kind "reference").
"""
import torch
import torch.nn.functional as F

from . import dmc_oracle as O
from . import ref_loader as R


class ReferenceTrainer:
    def __init__(self, num_class, hp, *, gan=False, arch_d=None, state=None, segments=3):
        variant = 'dmcnet_GAN' if gan else 'dmcnet'
        kw = dict(base_model='resnet18', arch_estimator='DenseNetTiny', gen_flow_or_delta=1, use_databn=0)
        if gan:
            kw['arch_d'] = arch_d
        self.model = R.build_reference_model(variant, num_class, segments, 'mv', **kw)
        if state is not None:
            self.model.load_state_dict(state)
        self.model.train()
        self.hp, self.gan, self.S = hp, gan, segments
        groups = {'base_model': [], 'gen_flow_model': [], 'discriminator': []}
        mults = {'base_model': hp.lr_cls_mult, 'gen_flow_model': hp.lr_mse_mult, 'discriminator': hp.lr_d_mult}
        for key, value in dict(self.model.named_parameters()).items():          # train.py:121-132
            for tag in groups:
                if tag in key:
                    groups[tag].append({'params': value, 'lr': hp.lr * mults[tag],
                                        'weight_decay': hp.weight_decay * (0.0 if 'bias' in key else 1.0)})
        mk = lambda g: torch.optim.Adam(g, eps=hp.eps, betas=hp.betas)           # train.py:134-142
        self.opt_cls, self.opt_gf = mk(groups['base_model']), mk(groups['gen_flow_model'])
        self.opt_d = mk(groups['discriminator']) if gan else None
        self.iteration = 0

    def step(self, input_flow, input_mv, input_residual, target):
        hp, S, model = self.hp, self.S, self.model
        flow = input_flow.view((-1,) + tuple(input_mv.shape[-3:]))               # train.py:230
        out = {}
        if not self.gan:
            output, gen_flow = model(input_mv, input_residual)                   # :236
            output = output.view((-1, S) + tuple(output.shape[1:])).mean(1)     # :239-240
            loss_cls = F.cross_entropy(output, target)
            loss_mse = F.mse_loss(gen_flow, flow)
            loss = loss_cls * hp.lr_cls + loss_mse * hp.lr_mse                   # :248
            steppers = (self.opt_cls, self.opt_gf)
            out['loss_mse'] = float(loss_mse.detach())
        else:
            valid = torch.ones(target.shape[0] * S, dtype=torch.int64)          # GAN/train.py:253-256
            fake = torch.zeros_like(valid)
            if self.iteration % 2 == 0:                                          # D-step :261-302
                output, validity, gen_flow = model(input_mv, input_residual, flow)
                output = output.view((-1, S) + tuple(output.shape[1:])).mean(1)
                loss_cls = F.cross_entropy(output, target)
                loss = loss_cls * hp.lr_cls + F.cross_entropy(validity, torch.cat((fake, valid), 0)) * hp.lr_adv_d
                steppers = (self.opt_cls, self.opt_d)
            else:                                                                # G-step :331-371
                output, validity, gen_flow = model(input_mv, input_residual)
                output = output.view((-1, S) + tuple(output.shape[1:])).mean(1)
                loss_cls = F.cross_entropy(output, target)
                loss_mse = F.mse_loss(gen_flow, flow)
                loss = loss_cls * hp.lr_cls + F.cross_entropy(validity, valid) * hp.lr_adv_g + loss_mse * hp.lr_mse
                steppers = (self.opt_gf,)
                out['loss_mse'] = float(loss_mse.detach())
        for op in (self.opt_cls, self.opt_gf, self.opt_d):
            if op is not None:
                op.zero_grad(set_to_none=False)
        loss.backward()
        for op in steppers:
            op.step()
        self.iteration += 1
        out.update(loss=float(loss.detach()), loss_cls=float(loss_cls.detach()))
        return out
