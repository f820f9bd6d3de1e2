"""Drop-in mirror of the reference package `rectified_spaattn` (same module and symbol names); the bodies call
the hand-written sm_100a kernels in librsa_b200.so through rsa_b200."""
