"""ilswiss_b200 -- B200-native (sm_100a) replacement for ILSwiss's off-policy hot path:
HBM replay ring + fused SAC / TD3 / AdvIRL gradient-step engine behind the reference's own
ReplayBuffer / Trainer / AdvIRL interfaces.  See DESIGN.md and INTEGRATION.md."""
__version__ = "0.1.0"
