"""dfgan-b200: B200-native (sm_100a) implementation of the DF-GAN training-step hot path of
Nandan-Upadhyaya/GAN-codes (Deep_Fusion_GAN/src/generator, src/discriminator and the step math of
src/deep_fusion_gan/model.py).  Import name: ``gan_codes_b200`` (see gan_codes_b200.py at the repo root).

Reference-facing modules (same names / signatures as the reference's classes and functions):
  modules          Generator, Discriminator, ResidualBlockG/D, AffineBlock      (generator/, discriminator/)
  deep_fusion_gan  DeepFusionGAN, prepare_data                                  (deep_fusion_gan/model.py, objects/utils.py)
  text_encoder     RNNEncoder                                                   (text_encoder/model.py)
  dataset, augment DFGANDataset, create_loader, BatchImageTransform             (objects/dataset.py, utils.py)
  evaluation       compute_inception_score, compute_fid, compute_is_fid         (evaluation/metrics.py)
  damsm, sample, train   damsm_loss; generate_images / save_image / sample; fix_seed / train
Engine underneath:
  engine           DFGANTrainer (fused iteration, CUDA graph, data parallel), FlatState (flat Adam)
  peer             PeerGroup: gradient buffers mapped across the ranks of one box (CUDA IPC / NVLS multicast) and the
                   exchange kernels over them (all-reduce, in-switch all-reduce, reduce-scatter / all-gather)
  gnet, dnet, convops, kernels, fp32, checkpoint, _lib (ctypes binding of include/dfgan_b200.h)
"""
__version__ = "0.1.0"
