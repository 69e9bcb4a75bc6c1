#!/usr/bin/env python
"""Golden vector for the FASTA reader behind ``GenomicSignatures.calculate``: the SOURCE TEXT of ``readFasta``
(seqUtils.py:46-69; the module as a whole is Python-2-only syntax) executed with ``.iteritems()`` -> ``.items()`` and an
``open`` that behaves like Python 2's text mode on POSIX (lines end at '\\n' only, nothing translated).  The input has a
tab in a header, CRLF line ends, a repeated id, a header that starts with blanks and an unterminated last line.
Build container only; the output ``ref_fasta.json`` is committed."""
import gzip
import io
import json
import logging
import os
import sys

REF = '/root/reference/dbb'
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_greedy import cut   # noqa: E402

FASTA = ('>seq1\tdescr with tab more\nACGTACGTAC\nGGGTTTAAAC\n'
         '>seq2 CRLF record\r\nACGTAC\r\nTTGACA\r\n'
         '>seq1 duplicate id replaces the first record\nTTTTGGGGCCCCAAAA\n'
         '>  leading blanks give an empty id\nACGTTGCA\n'
         '>dup first copy\nAAAAAAAACCCC\n>other\nGGGGGGGG\n>dup second copy replaces the first\nCCCCGGGGTTTT\n'
         '>last unterminated\nACGTACGTAAGG')


def main():
    src = open(os.path.join(REF, 'seqUtils.py')).read().replace('\r\n', '\n').replace('\r', '\n').split('\n')
    text = '\n'.join(cut(src, 'def readFasta(', 0)).replace('.iteritems()', '.items()')
    ns = {'gzip': gzip, 'logging': logging, 'sys': sys, 'open': lambda p: io.open(p, 'r', newline='\n')}
    exec(text, ns)
    path = os.path.join(HERE, '_tmp_golden.fasta')
    with io.open(path, 'w', newline='') as f:
        f.write(FASTA)
    seqs = ns['readFasta'](path)
    os.remove(path)
    with open(os.path.join(HERE, 'ref_fasta.json'), 'w') as f:
        json.dump({'fasta': FASTA, 'records': seqs}, f, indent=1, sort_keys=True)
    print(seqs)


if __name__ == '__main__':
    main()
