"""Launch tapes and exchange sequence numbers (host logic of the peer-memory SyncBN path, no GPU): every recorded
exchange draws a fresh sequence number at record time and again on every replay, in launch order, and a launch that
also carries a dropout key keeps both patches."""
import hostemu
from mmhand_b200.kernels import KeyRef


class FakeWorld:
    def __init__(self):
        self.peer, self.seq, self.size = 0xBEEF, 0, 2

    def next_seq(self):
        self.seq += 1
        return self.seq


def test_sequence_numbers_follow_launch_order_across_replays():
    ops = hostemu.ops()
    w = FakeWorld()
    seen = []

    def fake(*args):
        seen.append(args[:2])
        return 0

    with ops.record() as tape:
        for _ in range(3):
            args, patch = ops._seq_args(w, ("payload",))
            ops._run(fake, args, patch)
    assert seen == [(0xBEEF, 1), (0xBEEF, 2), (0xBEEF, 3)]
    tape.replay(step=1)
    tape.replay(step=2)
    assert [s for _, s in seen] == list(range(1, 10))


def test_key_patch_and_sequence_patch_compose():
    ops = hostemu.ops()
    w = FakeWorld()

    class Struct:
        drop_key = 0

    st = Struct()
    key = KeyRef(seed=7, layer_id=3)
    calls = []

    def fake(*args):
        calls.append((args[1], st.drop_key))
        return 0

    with ops.record() as tape:
        kp = ops._key(key, st)
        args, patch = ops._peer_args(w, ("x",), kp)
        ops._run(fake, args, patch)
        args0, patch0 = ops._peer_args(None, ("x",), kp)        # single GPU: (NULL, 0) prefix, only the key patch
        assert args0[:2] == [None, 0] and patch0 is kp
    tape.replay(step=5)
    assert calls[0] == (1, key.resolve(0)) and calls[1] == (2, key.resolve(5))
