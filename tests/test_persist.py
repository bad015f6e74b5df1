"""Chain records in josefine's sled/bincode vocabulary (josefine_b200/persist.py; chain.rs:117-205).

Parity UNPINNED against josefine (no Rust toolchain): byte layouts are hand-derived from bincode 1.3's
default encoding of the reference's derives.
"""
import pytest

from josefine_b200 import abi, persist
from oracle.restated import RestatedCluster

CAPTURE = abi.F_CAPTURE_MESSAGES | abi.F_CAPTURE_FSM


def test_block_record_bytes():
    # Block{id: BlockId::new(2), next: BlockId::new(1), data: vec![0xAA, 0xBB]}  (chain.rs:86-91)
    assert persist.encode_block(2, 1, b"\xaa\xbb").hex() == (
        "0800000000000000" "0000000000000002"      # id: u64 len 8, then the big-endian id bytes
        "0800000000000000" "0000000000000001"      # next
        "0200000000000000" "aabb")                 # data: u64 len 2, bytes
    # the genesis record Chain::init writes (chain.rs:139-153): id 0, next 0, no data
    assert persist.encode_block(0, 0, b"").hex() == "0800000000000000" + "00" * 8 + "0800000000000000" + "00" * 8 + "00" * 8
    assert persist.decode_block(persist.encode_block(7, 5, b"xyz")) == (7, 5, b"xyz")
    for bad in (b"", b"\x08" + b"\x00" * 7, persist.encode_block(1, 0, b"a") + b"\x00", persist.encode_block(1, 0, b"ab")[:-1]):
        with pytest.raises(ValueError):
            persist.decode_block(bad)


def test_reference_block_id_serde():
    # chain.rs:345-350: bincode round trip of BlockId::new(0)
    rec = persist.encode_block(0, 0, b"")
    assert rec[:16].hex() == "0800000000000000" + "00" * 8          # bincode(BlockId(0)): u64 length 8 + the 8 id bytes
    assert persist.decode_block(rec)[0] == 0


def test_commit_key_sorts_inside_the_block_keyspace():
    # D6: b"commit" = 63 6f 6d 6d 69 74 lies between block ids 0x636f6d6d6973ffff.. and 0x636f6d6d69740000..
    assert persist.block_key(1) < persist.COMMIT_KEY < persist.block_key(0x64 << 56)
    assert persist.block_key(0x636F6D6D69740000) > persist.COMMIT_KEY > persist.block_key(0x636F6D6D6973FFFF)


def committed_cluster(make):
    api = make(1, 3, flags=CAPTURE, seed=5)
    now, payloads = 0, {}
    for step in range(70):
        now += 100
        prop = None
        if step >= 30 and step % 4 == 0:
            tok = 1000 + step
            payloads[tok] = b"value-%d" % step
            prop = [(1 + step % 3, tok)]
        api.step(now, proposals=prop)
    return api, payloads


def check_records(api, payloads):
    for node in (1, 2, 3):
        st = api.query(0, node)
        recs = persist.chain_records(api, 0, node, payloads)
        assert recs == sorted(recs) and len({k for k, _ in recs}) == len(recs)
        assert recs[0] == (b"\x00" * 8, persist.encode_block(0, 0, b""))            # genesis, chain.rs:139-153
        tree = persist.reopen(recs)
        assert tree["commit"] == st.commit == tree["head"] == tree["id_gen"]         # chain.rs:125-130
        assert st.commit > 0 and recs[-1] == (persist.COMMIT_KEY, persist.block_key(st.commit))
        live = [b for b in api.chain_read(0, node, 0, st.max_key + 1) if b is not None]
        assert sorted(tree["blocks"]) == [b[0] for b in live]
        for bid, nxt, tok in live:
            assert tree["blocks"][bid] == (nxt, payloads.get(tok, b""))
        # walking `next` from the commit reaches genesis: the committed branch is whole
        at, hops = st.commit, 0
        while at != 0:
            at, hops = tree["blocks"][at][0], hops + 1
        assert hops >= 1


def test_records_of_a_committed_group():
    api, payloads = committed_cluster(RestatedCluster.create)
    check_records(api, payloads)
    api.compact()                                                                    # chain.rs:239-253
    check_records(api, payloads)


def test_no_commit_key_before_the_first_commit():
    api = RestatedCluster.create(1, 3, flags=CAPTURE)
    recs = persist.chain_records(api, 0, 1)
    assert recs == [(b"\x00" * 8, persist.encode_block(0, 0, b""))]
    assert persist.reopen(recs)["commit"] == 0
    with pytest.raises(ValueError):
        persist.reopen([(persist.COMMIT_KEY, b"\x01")])
    with pytest.raises(ValueError):
        persist.reopen([(persist.block_key(3), persist.encode_block(4, 0, b""))])


@pytest.mark.gpu
def test_records_from_the_engine_match_the_oracle():
    from josefine_b200 import RaftEngine
    eng, payloads = committed_cluster(RaftEngine.create)
    ora, _ = committed_cluster(RestatedCluster.create)
    check_records(eng, payloads)
    for node in (1, 2, 3):
        assert persist.chain_records(eng, 0, node, payloads) == persist.chain_records(ora, 0, node, payloads)


def _restart_roundtrip(make):
    """Export a follower's chain as sled records, restart the node from them (jr_node_restart = Chain::new over a
    persisted tree, chain.rs:117-137): head = commit = id_gen = the persisted commit, blocks intact."""
    api, payloads = committed_cluster(make)
    st = api.query(0, 2)
    recs = persist.chain_records(api, 0, 2, payloads)
    before = api.chain_read(0, 2, 0, int(st.max_key) + 1)
    tokens = {v: k for k, v in payloads.items()}
    persist.restart_from_records(api, 0, 2, 9000, recs, tokens)
    after = api.query(0, 2)
    assert (after.head, after.commit, after.id_gen) == (st.commit, st.commit, st.commit) and st.commit > 0
    assert (after.current_term, after.voted_for, after.role, after.fault) == (0, 0, abi.ROLE_FOLLOWER, 0)
    assert api.chain_read(0, 2, 0, int(st.max_key) + 1) == before
    return api


def test_restart_from_records_on_oracle():
    _restart_roundtrip(lambda g, r, **kw: RestatedCluster.create(g, r, **kw))


def test_restart_from_records_on_device_code():
    from tests.emu.emu import EmuEngine
    _restart_roundtrip(lambda g, r, **kw: EmuEngine.create(g, r, **kw))


@pytest.mark.gpu
def test_restart_from_records_on_gpu():
    from josefine_b200 import RaftEngine
    _restart_roundtrip(lambda g, r, **kw: RaftEngine.create(g, r, **kw))
