"""Policy plugin surface — drop-in for ``visual_mpc/policy/policy.py`` (reference lines 9-81).

``Policy`` subclasses are constructed as ``cls(ag_params, policyparams, gpu_id, ngpu)``
(``sim/simulator.py:21``), ``reset()`` once per trajectory (``agent/general_agent.py:193``) and
``act(**get_policy_args(...))`` once per step (``general_agent.py:206``)."""
from __future__ import annotations

import abc
import inspect
from typing import Any, Dict, Optional

import numpy as np

from .hparams import HParams

_SPECIAL = ("t", "i_tr", "obs", "step_data", "goal_pos")


def get_policy_args(policy, obs: Dict[str, Any], t: int, i_tr: int, step_data: Optional[Dict[str, Any]] = None):
    """Resolve ``policy.act``'s keyword arguments BY NAME: observation dict first, then the agent's
    step data, then the specials (t, i_tr, obs, step_data, goal_pos); a parameter with neither a
    value nor a default is an error (reference policy.py:9-46)."""
    resolved = {}
    for name, param in inspect.signature(policy.act).parameters.items():
        if param.kind in (param.VAR_POSITIONAL, param.VAR_KEYWORD):
            continue
        if name in obs:
            val = obs[name]
        elif step_data is not None and name in step_data:
            val = step_data[name]
        elif name in _SPECIAL:
            val = {"t": t, "i_tr": i_tr, "obs": obs, "step_data": step_data}.get(name) if name != "goal_pos" \
                else step_data["goal_pos"]
        else:
            val = param.default
        if val is inspect.Parameter.empty:
            raise ValueError("Required Policy Param {} not set in agent".format(name))
        resolved[name] = val
    return resolved


class Policy(abc.ABC):
    """Base class.  Subclasses populate ``self._hp`` from ``_default_hparams()`` then call
    ``_override_defaults(policyparams)``."""

    def _default_hparams(self) -> HParams:
        return HParams()

    def _override_defaults(self, policyparams: Dict[str, Any]) -> None:
        """Unknown keys fail; an override EQUAL to the default raises (reference policy.py:51-63 —
        shipped configs rely on this never firing)."""
        for key, val in policyparams.items():
            if key == "type":                    # the policy class itself
                continue
            if key not in self._hp:
                raise AttributeError("unknown policy hyper-parameter %r" % key)
            current = self._hp.get(key)
            same = False
            try:
                same = bool(np.all(val == current))
            except Exception:
                same = False
            if same:
                raise ValueError("attribute is {} is identical to default value!!".format(key))
            if current is None:
                setattr(self._hp, key, val)      # no type check against a None default
            else:
                self._hp.set_hparam(key, val)

    @abc.abstractmethod
    def act(self, *args, **kwargs) -> Dict[str, Any]:
        """Returns a dict whose 'actions' entry is the (adim,) action for this step."""

    def reset(self) -> None:
        pass


class NullPolicy(Policy):
    """Always returns the zero action (reference policy.py:97-117)."""

    def __init__(self, ag_params, policyparams, gpu_id=0, ngpu=1):
        self._adim = ag_params["adim"]
        self._hp = self._default_hparams()
        self._override_defaults(policyparams)

    def _default_hparams(self):
        hp = super()._default_hparams()
        hp.add_hparam("wait_for_user", False)
        return hp

    def act(self):
        return {"actions": np.zeros(self._adim)}
