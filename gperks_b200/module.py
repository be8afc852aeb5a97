"""Base class giving torch modules the gpytorch ``Module`` conveniences GPErks relies on:
``register_constraint``, dotted-name ``initialize(**{"a.b.raw_x": t})`` and constrained-value setters
(call sites: GPErks/train/emulator.py:54-58, 240-244)."""
from __future__ import annotations

from typing import Iterator, Tuple

import torch
from torch import nn

PARAM_DTYPE = torch.float64   # the B200 path computes in FP64; parameters are kept in FP64


class Module(nn.Module):
    def register_constraint(self, param_name: str, constraint, replace: bool = True) -> None:
        if param_name not in self._parameters:
            raise RuntimeError(f"Attempting to register constraint for nonexistent parameter {param_name!r}.")
        name = param_name + "_constraint"
        if name in self._modules and not replace:
            return
        self.add_module(name, constraint)

    def constraint_for_parameter_name(self, param_name: str):
        mod = self
        *path, leaf = param_name.split(".")
        for part in path:
            mod = getattr(mod, part)
        return mod._modules.get(leaf + "_constraint")

    def named_hyperparameters(self) -> Iterator[Tuple[str, nn.Parameter]]:
        return self.named_parameters()

    def hyperparameters(self):
        return self.parameters()

    def named_parameters_and_constraints(self):
        for name, p in self.named_parameters():
            yield name, p, self.constraint_for_parameter_name(name)

    def initialize(self, **kwargs):
        """Set parameters (or constrained properties) by name; names may be dotted paths."""
        for name, val in kwargs.items():
            if isinstance(val, int):
                val = float(val)
            if "." in name:
                head, rest = name.split(".", 1)
                sub = getattr(self, head)
                if not isinstance(sub, Module):
                    raise AttributeError(f"{head} is not a gperks_b200 Module")
                sub.initialize(**{rest: val})
                continue
            if name in self._parameters:
                p = self._parameters[name]
                with torch.no_grad():
                    if torch.is_tensor(val):
                        v = val.detach().to(p.device, p.dtype)
                        if v.numel() == p.numel():
                            p.copy_(v.reshape(p.shape))
                        else:
                            p.copy_(v.expand_as(p))
                    else:
                        p.fill_(val)
            elif isinstance(getattr(type(self), name, None), property):
                setattr(self, name, val)
            elif hasattr(self, name):
                setattr(self, name, val)
            else:
                raise AttributeError(f"Unknown parameter {name} for {type(self).__name__}")
        return self

    # helpers for subclasses -----------------------------------------------------------------
    def _constrained(self, raw_name: str) -> torch.Tensor:
        raw = self._parameters[raw_name]
        c = self._modules.get(raw_name + "_constraint")
        return raw if c is None else c.transform(raw)

    def _set_constrained(self, raw_name: str, value) -> None:
        raw = self._parameters[raw_name]
        if not torch.is_tensor(value):
            value = torch.as_tensor(value, dtype=raw.dtype, device=raw.device)
        value = value.to(raw.device, raw.dtype)
        c = self._modules.get(raw_name + "_constraint")
        self.initialize(**{raw_name: value if c is None else c.inverse_transform(value)})
