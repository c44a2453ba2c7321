"""kccotgan_b200 — B200-native causal-OT loss path of KCCOT-GAN (drop-in for the reference's
`gan_utils` functions and `data_utils.KernelSmoothing`).  See DESIGN.md."""
__version__ = "0.1.0"
