"""Host helpers the hot path uses from the reference's `glow/utils.py`."""


def get_longest_history(cond_params):
    """start_ts of the frame loop: the longest conditioning window (utils.py:44-50)."""
    return max(cond_params[m]["history"] for m in ("p1_face", "p1_speech", "p2_speech", "p2_face"))


def calc_jerk(x):
    """Mean absolute third difference along time of [B, T, C] (utils.py:53-58)."""
    x = x.cpu()
    d1 = x[:, 1:] - x[:, :-1]
    d2 = d1[:, 1:] - d1[:, :-1]
    d3 = d2[:, 1:] - d2[:, :-1]
    return d3.abs().mean()


def test_params(hparams):
    """History must be shorter than the train / validation sequence length (utils.py:116-122)."""
    longest = get_longest_history(hparams.Conditioning)
    for split in ("Train", "Validation"):
        assert getattr(hparams, split)["seq_len"] > longest, "Sequence length (%s) must be longer than the history" % split


def derange_batch(batch_data, modalities, shuffle_time=False):
    """Shuffles the conditioning modalities across the batch (utils.py:85-100): the mismatched-NLL probe of
    `LetsFaceItGlow.training_step`.  RNG consumption as the reference: one `torch.randperm(batch_size)` on the CPU
    generator (plus one per modality over time with `shuffle_time`); the gathers run on the tensors' own device."""
    import torch

    batch_size = batch_data["p1_face"].size(0)
    permutation = torch.randperm(batch_size)
    mixed_up_batch = {}
    for modality in ["p1_face", "p2_face", "p1_speech", "p2_speech"]:
        if modality in modalities:
            x = batch_data[modality]
            mixed_up_batch[modality] = x[permutation.to(x.device)]
            if shuffle_time:
                t_perm = torch.randperm(x.size(1))
                mixed_up_batch[modality] = mixed_up_batch[modality][:, t_perm.to(x.device)]
        elif batch_data.get(modality) is not None:
            mixed_up_batch[modality] = batch_data[modality]
    return mixed_up_batch


def get_mismatched_modalities(hparams):
    """Which interlocutor modalities the mismatched-NLL probe deranges, and its log name (utils.py:103-113)."""
    modalities = []
    if hparams.Conditioning["p2_face"]["history"] > 0:
        modalities.append("p2_face")
    if hparams.Conditioning["p2_speech"]["history"] > 0:
        modalities.append("p2_speech")
    name = "p2" if len(modalities) == 2 else modalities[0]
    return modalities, name
