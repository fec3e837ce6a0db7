"""CPU: the consensus oracle reproduces the reference's known-answer test
(tests/preprocessing/test_consensus.py:13-80 of the reference)."""


def known_answer_reads():
    from oracle.consensus_oracle import make_read
    return [
        make_read('read6', 12, [('M', 2)], 'GG', [20] * 2, '1C'),
        make_read('read1', 1, [('M', 4)], 'ACTG', [30] * 4, '4'),
        make_read('read3', 5, [('M', 4)], 'ACAG', [30, 40, 30, 30], '0C0T2'),
        make_read('read2', 3, [('M', 4)], 'TGCT', [30] * 4, '4'),
        make_read('read5', 9, [('M', 2), ('D', 1), ('M', 1)], 'CTG', [30] * 3, '2^A1'),
        make_read('read4', 7, [('M', 4), ('I', 1)], 'AGCTA', [30] * 5, '4'),
    ]


def test_known_answer():
    from oracle.consensus_oracle import call_consensus
    r = call_consensus(known_answer_reads(), quality=27)
    assert r['seq'] == 'ACTGCCAGCTGC'
    assert r['start'] == 1 and r['end'] == 14
    assert ''.join(f'{n}{op}' for op, n in r['cigar']) == '10M1D2M'
    assert r['md'] == '5T4^A2'
    assert r['nm'] == 1
