"""Host helpers the hot path uses from the reference's `glow/utils.py`."""


def get_longest_history(cond_params):
    """start_ts of the frame loop: the longest conditioning window (utils.py:44-50)."""
    return max(cond_params[m]["history"] for m in ("p1_face", "p1_speech", "p2_speech", "p2_face"))


def calc_jerk(x):
    """Mean absolute third difference along time of [B, T, C] (utils.py:53-58), the validation metric of
    mimicry_logger.py:175-184.  The reference moves the frames to the host first (`x.cpu()`); here the generated frames stay
    where `SeqGlow.inference` left them: one launch (`lfi_jerk`) and a one-float result on the device.  CUDA tensors only."""
    import torch

    from .. import _cabi as cabi

    if x.device.type != "cuda":
        raise RuntimeError("lets_face_it_b200.calc_jerk: frames live on %s; the path runs on CUDA only (no CPU path)" % x.device)
    if x.dim() != 3 or x.shape[1] < 4:
        raise RuntimeError("calc_jerk expects [B, T >= 4, C] frames, got %s" % (tuple(x.shape),))
    x = x.detach().to(torch.float32).contiguous()
    out = torch.empty(1, dtype=torch.float32, device=x.device)
    scratch = torch.empty(1, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        cabi.check(cabi.lib().lfi_jerk(x.data_ptr(), x.shape[0], x.shape[1], x.shape[2], scratch.data_ptr(), out.data_ptr(),
                                       cabi.stream_ptr()), "lfi_jerk")
    return out[0]


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
