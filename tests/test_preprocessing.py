"""The steps in front of the hot path, checked against the reference's OWN code
and files where they exist (build container) and against known answers everywhere:

* rgcn_b200.ntriples: vectorised N-Triples -> node / relation ids -> edge tensors == Graph.init_graph
  (/root/reference/graphs/graph.py:24-69) on every shipped graph; map index == the dict walk of
  model/embeddingTricks.py:17-23 over graphs/graphProcessing.py:41-52's dicts;
* rgcn_b200.attribute_summary: MurmurHash3-x64-128 == mmh3's published values and == the node ids inside the
  reference's shipped AIFB summary files; create_sum_map == the reference's create_sum_map byte for byte.
"""
import glob
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ORACLE, REFERENCE, load_golden

from rgcn_b200 import attribute_summary, ntriples

HAVE_REF = os.path.isdir(REFERENCE)
needs_ref = pytest.mark.skipif(not HAVE_REF, reason='needs /root/reference (build container only)')


def test_murmur3_known_answers():
    h = attribute_summary.hash128
    assert h('foo') == 168394135621993849475852668931176482145            # mmh3.hash128('foo')
    assert (h('foo') & ((1 << 64) - 1)) - (1 << 64) == -2129773440516405919 and h('foo') >> 64 == 9128664383759220103   # mmh3.hash64
    assert h('') == 0
    assert h(b'foo') == h('foo') and h('foo', seed=1) != h('foo')
    # every tail length of the 16-byte block loop
    seen = {h('a' * n) for n in range(0, 40)}
    assert len(seen) == 40


def test_parse_quirks_and_inverse_edges():
    lines = ['<B> <p> <a> .', '', '<a> <http://www.w3.org/1999/02/22-rdf-syntax-ns#type> <C> .', '<a> <Q> "Lit x" .',
             '<B> <p> <a> .', '<z>  <p> <a>  .']     # duplicate kept; double space -> empty predicate / trailing space kept
    g = ntriples.parse_graph(lines)
    assert g.nodes == sorted({'<b>', '<a>', '<c>', '"lit x"', '<z>', '<p> <a> '})
    assert g.relations == {'': 0, '<p>': 1, '<q>': 2}
    ei, et = g.edge_index, g.edge_type
    assert not ei.is_contiguous() and ei.shape == (2, 8) and et.tolist() == [2, 3, 4, 5, 2, 3, 0, 1]
    n = g.node_to_enum
    assert ei[:, 0].tolist() == [n['<b>'], n['<a>']] and ei[:, 1].tolist() == [n['<a>'], n['<b>']]
    assert g.num_edges == 5
    with pytest.raises(IndexError):
        ntriples.parse_graph(['<a> <b> .'])


def _reference_graph(path):
    sys.path.insert(0, os.path.join(ORACLE, 'shim'))
    sys.path.insert(0, ORACLE)
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    from graphs.graph import Graph
    from graphs.graphProcessing import parse_graph_nt
    g = Graph(os.path.basename(path), {})
    g.init_graph(parse_graph_nt(path))
    return g


@needs_ref
@pytest.mark.parametrize('rel_path', ['graphs/TEST/TEST_complete.nt', 'graphs/TEST/attr/sum/TEST_sum_in_out.nt',
                                      'graphs/AIFB/attr/sum/AIFB_sum_in.nt', 'graphs/AIFB/attr/sum/AIFB_sum_out.nt',
                                      'graphs/AIFB/bisim/sum/AIFB_bisim_k3.nt', 'graphs/AIFB/dummy/sum/AIFB_sum_dummy.nt',
                                      'graphs/MUTAG/bisim/sum/MUTAG_bisim_k1.nt'])
def test_parse_graph_equals_reference_init_graph(rel_path):
    path = os.path.join(REFERENCE, rel_path)
    if not os.path.exists(path):
        pytest.skip('file not shipped')
    ref = _reference_graph(path)
    got = ntriples.parse_graph_file(path, relation_order=ref.relations)
    assert got.nodes == ref.nodes and got.node_to_enum == ref.node_to_enum and got.num_nodes == ref.num_nodes
    assert got.relations == ref.relations and got.num_edges == ref.num_edges
    td = ref.training_data
    assert torch.equal(got.edge_index, td.edge_index) and torch.equal(got.edge_type, td.edge_type)
    assert got.edge_index.stride() == td.edge_index.stride() and got.edge_type.stride() == td.edge_type.stride()
    # canonical relation ids: a permutation of the reference's, same edges
    canon = ntriples.parse_graph_file(path)
    perm = torch.tensor([canon.relations[k] for k, _ in sorted(ref.relations.items(), key=lambda kv: kv[1])])
    assert torch.equal(canon.edge_index, td.edge_index)
    assert torch.equal(canon.edge_type, 2 * perm[td.edge_type // 2] + td.edge_type % 2)


def test_parse_graph_matches_golden_fixture_edges():
    """Without the reference: the golden edge tensors (made by the reference) are internally consistent with the
    parser's emission rule — every odd column is the inverse of the even one with relation id + 1."""
    g = load_golden('graph_AIFB_bisim_k3.npz')
    ei, et = g['edge_index'], g['edge_type']
    assert (ei[0, 0::2] == ei[1, 1::2]).all() and (ei[1, 0::2] == ei[0, 1::2]).all() and (et[1::2] == et[0::2] + 1).all()


@needs_ref
@pytest.mark.parametrize('tag', ['AIFB/attr', 'AIFB/bisim', 'TEST/attr'])
def test_map_index_equals_reference_dict_walk(tag):
    sys.path.insert(0, ORACLE)
    import rgcn_oracle
    sums = sorted(glob.glob(os.path.join(REFERENCE, 'graphs', tag, 'sum', '*.nt')))
    maps = sorted(glob.glob(os.path.join(REFERENCE, 'graphs', tag, 'map', '*.nt')))
    assert sums and len(sums) == len(maps)
    for sp, mp in list(zip(sums, maps))[:3]:
        sg = ntriples.parse_graph_file(sp)
        lines = ntriples.read_lines(mp)
        org2sum, sum2org = ntriples.node_mappings(lines)
        from graphs.graphProcessing import get_node_mappings_dict, parse_graph_nt
        r_org2sum, r_sum2org = get_node_mappings_dict(parse_graph_nt(mp))
        assert org2sum == dict(r_org2sum) and sum2org == dict(r_sum2org)
        assert list(org2sum) == list(r_org2sum)                                # same (sorted) key order
        org_nodes = sorted(org2sum)[::1]
        if len(org_nodes) > 5:
            org_nodes = org_nodes[:-3] + ['<not-in-the-map>']                   # unmapped original nodes -> -1
        want = rgcn_oracle.build_map_index({n: i for i, n in enumerate(org_nodes)}, sg.node_to_enum, org2sum)
        got = ntriples.map_index(lines, org_nodes, sg.node_to_enum)
        assert got.dtype == torch.int32 and torch.equal(got, want.to(torch.int32))


@needs_ref
def test_hash_reproduces_the_node_ids_of_the_shipped_aifb_summaries():
    """AIFB_sum_out.nt (written by the reference with the real mmh3): a summary node's id must be the hash of the
    sorted, comma-joined set of non-type predicates on its outgoing lines; same for incoming on AIFB_sum_in.nt."""
    for fname, col in (('AIFB_sum_out.nt', 0), ('AIFB_sum_in.nt', 2)):
        s, p, o = ntriples.tokenize(ntriples.read_lines(os.path.join(REFERENCE, 'graphs/AIFB/attr/sum', fname)))
        node = (s, p, o)[col]
        sets = {}
        for nd, pr in zip(node.tolist(), p.tolist()):
            if pr != attribute_summary.RDF_TYPE:
                sets.setdefault(nd, set()).add(pr)
        checked = 0
        for nd, ps in sets.items():
            if nd == '<0>':
                continue
            assert nd == f'<{attribute_summary.hash128(",".join(sorted(ps)))}>', (fname, nd)
            checked += 1
        assert checked > 20


@needs_ref
def test_create_sum_map_equals_reference_byte_for_byte(tmp_path):
    """The reference's createAttributeSum.create_sum_map, run with this package's hash128 standing in for the absent
    mmh3 module, against rgcn_b200.attribute_summary.create_sum_map on the shipped TEST graph and on a synthetic one
    with literals, duplicates and type triples."""
    fake = types.ModuleType('mmh3')
    fake.hash128 = attribute_summary.hash128
    saved = sys.modules.get('mmh3')
    sys.modules['mmh3'] = fake
    try:
        if REFERENCE not in sys.path:
            sys.path.insert(0, REFERENCE)
        sys.modules.pop('graphs.createAttributeSum', None)
        ref_mod = importlib.import_module('graphs.createAttributeSum')
        rng = np.random.default_rng(0)
        synth = tmp_path / 'SYN_complete.nt'
        with open(synth, 'w') as f:
            for _ in range(400):
                s_, p_, o_ = rng.integers(0, 40), rng.integers(0, 7), rng.integers(0, 60)
                obj = f'"Literal {o_}"' if o_ % 5 == 0 else f'<http://x/N{o_}>'
                pred = attribute_summary.RDF_TYPE if p_ == 0 else f'<http://x/P{p_}>'
                f.write(f'<http://x/N{s_}> {pred} {obj} .\n')
        for src, name in ((os.path.join(REFERENCE, 'graphs/TEST/TEST_complete.nt'), 'TEST'), (str(synth), 'SYN')):
            for who, fn in (('ref', ref_mod.create_sum_map), ('mine', attribute_summary.create_sum_map)):
                d = tmp_path / f'{name}_{who}'
                os.makedirs(d / 'sum')
                os.makedirs(d / 'map')
                fn(src, str(d / 'sum') + '/', str(d / 'map') + '/', name)
            for sub in ('sum', 'map'):
                for f_ref in sorted(glob.glob(str(tmp_path / f'{name}_ref' / sub / '*.nt'))):
                    f_mine = f_ref.replace(f'{name}_ref', f'{name}_mine')
                    assert open(f_ref).read() == open(f_mine).read(), f_ref
    finally:
        if saved is not None:
            sys.modules['mmh3'] = saved
        else:
            sys.modules.pop('mmh3', None)
        sys.modules.pop('graphs.createAttributeSum', None)
