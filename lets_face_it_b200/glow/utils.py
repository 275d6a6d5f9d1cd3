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
