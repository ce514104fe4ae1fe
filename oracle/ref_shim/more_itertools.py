"""Stand-in for the two `more_itertools` helpers the reference imports."""
import itertools


def distinct_permutations(iterable, r=None):
    """Distinct permutations in sorted order (what more_itertools yields for sortable input)."""
    return iter(sorted(set(itertools.permutations(sorted(iterable), r))))


def consume(iterator, n=None):
    for _ in (iterator if n is None else itertools.islice(iterator, n)):
        pass
