"""humanoid_amp_b200 -- the per-step data-parallel AMP path of zhoushanghai/humanoid_amp on B200 (sm_100a).

Public surface (same names and argument meaning as the reference, see each module's docstring for file:line):

* ``MotionLoader``                                  <- ``motions/motion_loader.py``
* ``compute_obs``, ``quaternion_to_tangent_and_normal``, ``AmpEnvPath`` (``collect_reference_motions``, history update,
  reset fill)                                        <- ``g1_amp_env.py`` / ``humanoid_amp_env.py``
* ``AmpDiscriminator.style_reward``                  <- skrl ``AMP._update`` style-reward block
* ``reduce_parameters`` / ``GradientBucket`` / ``shard_envs``  <- skrl ``Model.reduce_parameters`` (NVLink peer-memory all-reduce) + env sharding
* ``AmpStateMemory``, ``RunningStandardScaler``        <- skrl ``RandomMemory`` (motion dataset / replay buffer), AMP state preprocessor
* ``AmpDiscriminatorUpdate``                         <- skrl ``AMP._update`` discriminator loss block + its backward pass
* ``InputPrefetcher`` / ``ResultReader``             host staging: the reference's host ``times`` / ``motion_ids`` in, rewards out

Everything executes in ``libamp_b200.so`` (hand-written CUDA behind the C ABI of ``include/amp_b200.h``); importing
this package needs no GPU, calling it does, and there is no CPU fallback.
"""

from ._lib import AmpB200Error  # noqa: F401
from .robots import G1, HUMANOID28, RobotSpec, robot_for_clip  # noqa: F401
from .motion_loader import MotionLoader, _resolve_motion_files  # noqa: F401
from .amp_env import AmpEnvCfg, AmpEnvPath, compute_obs, quaternion_to_tangent_and_normal  # noqa: F401
from .discriminator import AmpDiscriminator, style_reward_from_logits  # noqa: F401
from .distributed import GradientBucket, reduce_parameters, shard_envs  # noqa: F401
from .graphs import capture_step  # noqa: F401
from .pipeline import InputPrefetcher, ResultReader  # noqa: F401
from .memory import AmpStateMemory, RunningStandardScaler  # noqa: F401
from .disc_update import AmpDiscriminatorUpdate  # noqa: F401

__version__ = "0.1.0"
