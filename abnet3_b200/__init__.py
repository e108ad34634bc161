"""abnet3_b200 -- the data-parallel hot path of bootphon/abnet3 on B200 (sm_100a).

Aligning same-word token pairs (cosine frame distance -> DTW -> traceback ->
aligned frame-index pairs) and training the siamese frame embedder on them
(MLP forward/backward, coscos2 / cosine-margin loss, optimizer step), behind
the reference's own Python surface:

    abnet3_b200.utils       cosine_distance, DTW, get_dtw_alignment, Features_Accessor, ...
    abnet3_b200.dataloader  OriginalDataLoader, FramesDataLoader, MultiTaskDataLoader,
                            PairsDataLoader
    abnet3_b200.model       SiameseNetwork, SiameseMultitaskNetwork
    abnet3_b200.loss        coscos2, cosmargin, weighted_loss_multi
    abnet3_b200.trainer     TrainerSiamese, TrainerSiameseMultitask (+ NCCL data parallelism)

All arithmetic runs in hand-written CUDA kernels behind a C ABI
(include/abnet3_b200.h, abnet3_b200/libabnet3_b200.so).  There is no CPU
fallback: importing the kernels' wrappers without the built library, or
calling them without an sm_100 GPU, raises.
"""
__version__ = "0.1.0"
