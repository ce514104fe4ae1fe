"""Stand-in for `collections_extended.bijection` (dict with an `inverse` view)."""


class bijection(dict):
    @property
    def inverse(self):
        return {v: k for k, v in self.items()}
