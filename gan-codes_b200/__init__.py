"""dfgan-b200: B200-native (sm_100a) implementation of the DF-GAN training-step hot path of
Nandan-Upadhyaya/GAN-codes (Deep_Fusion_GAN/src/generator, src/discriminator and the step math of
src/deep_fusion_gan/model.py).  Import name: ``gan_codes_b200`` (see gan_codes_b200.py at the repo root).
"""
__version__ = "0.1.0"
