"""CPU oracle for the AMP hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This package restates, in numpy + CPU torch (fp32, float64 host index math), the algorithm of the
reference's per-step AMP path so the CUDA kernels in ``humanoid_amp_b200`` can be checked against it:

* ``motion_oracle``  -- ``MotionLoader`` (reference ``motions/motion_loader.py:98-390``)
* ``env_oracle``     -- ``compute_obs`` / ``quaternion_to_tangent_and_normal`` / ``collect_reference_motions`` /
                        history shift / reset fill (reference ``g1_amp_env.py:175-193, 414-419, 445-497, 535-561``)
* ``memory_oracle``  -- skrl ``RandomMemory`` ring write / ``sample_by_index`` (upstream skrl, PARITY UNPINNED)
* ``disc_train_oracle`` -- skrl ``AMP._update`` discriminator LOSS block (BCE + logit regularisation + gradient penalty + weight decay)
                        evaluated by torch autograd, and the closed-form gradients the CUDA path implements (PARITY UNPINNED)
* ``dataset_oracle`` -- the offline dataset tool ``motions/data_convert.py:161-379``: scipy interpolation, restated
                        Pinocchio forward kinematics / Eigen quaternion conversion, velocity stages
* ``disc_oracle``    -- skrl ``RunningStandardScaler`` (eval + train-mode statistics) + MLP + AMP style reward (upstream skrl >= 1.4.3,
                        ``agents/torch/amp/amp.py::_update``; configured by ``agents/skrl_g1_dance_amp_cfg.yaml:31-39, 80, 94-95``)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this package, and only as the checker or the timed CPU baseline.  The product package never imports it and has no CPU
fallback: without the CUDA extension it raises.

Parity pinning status
---------------------
* ``motion_oracle``: PINNED -- bit-identical to the live reference ``MotionLoader`` (imported by file path from
  ``/root/reference``) on all 8 shipped clips; see ``tests/golden/make_golden.py`` and ``tests/test_oracle_pins.py``.
  The committed fixtures under ``tests/golden/`` were produced by the live reference, not by this oracle.
* ``env_oracle``: PINNED ON THE REFERENCE'S OWN TEXT -- the env modules import ``isaaclab`` (absent) and cannot be imported,
  so ``oracle/build_ref.py`` cuts the source of the methods on the path (``_get_observations``, ``_get_rewards``,
  ``_reset_strategy_random``, ``collect_reference_motions``, ``compute_obs``, ``quaternion_to_tangent_and_normal``,
  ``compute_rewards``, ``exp_reward_with_floor``) out of ``g1_amp_env.py`` / ``humanoid_amp_env.py`` unmodified and
  ``oracle/ref_harness.py`` executes it on CPU torch; ``tests/golden/make_golden.py`` wrote every env fixture with that text and
  ``env_oracle`` must reproduce them bit for bit (``tests/test_oracle_pins.py``).  Only ``quat_apply`` / ``quat_rotate_inverse``
  (upstream Isaac Lab 2.2.0 ``isaaclab.utils.math``, not vendored) are injected restatements: PARITY UNPINNED for those two
  functions, convention checked against scipy's ``Rotation``.
* ``dataset_oracle``: PINNED -- ``motions/data_convert.py`` run UNMODIFIED (Pinocchio replaced by the restated forward
  kinematics) wrote ``tests/golden/dataset/data_convert_output.npz``, which the oracle reproduces bit for bit; the restated FK is
  pinned on ``motions/custom_motion.npz``, the reference's own output of that conversion made with the real Pinocchio (poses to
  2 float32 ulp; ``tests/test_dataset_oracle.py``).
* ``disc_oracle``: skrl is a third-party dependency that is neither vendored nor installed: PARITY UNPINNED; the oracle
  is a literal restatement of the upstream expression.
"""

from .motion_oracle import OracleMotionLoader, frame_blend_f64, lerp_f32, slerp_f32  # noqa: F401
from .env_oracle import (  # noqa: F401
    quat_apply,
    quaternion_to_tangent_and_normal,
    compute_obs,
    collect_reference_motions,
    history_times,
    shift_and_write_history,
    reset_fill,
    actor_observations,
    reset_root_and_dof_state,
    quat_rotate_inverse,
    exp_reward_with_floor,
    task_rewards,
)
from .disc_oracle import OracleDiscriminator, running_standard_scaler_eval, style_reward_from_logits  # noqa: F401
from .memory_oracle import OracleRandomMemory  # noqa: F401
from .disc_train_oracle import DiscLossCfg, discriminator_loss_autograd, discriminator_loss_manual  # noqa: F401
