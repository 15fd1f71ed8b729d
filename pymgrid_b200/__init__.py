"""pymgrid_b200 -- B200-native batched microgrid-step engine behind pymgrid's Microgrid / envs surface."""
__version__ = "0.1.0"
