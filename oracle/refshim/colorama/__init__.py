class _C:
    def __getattr__(self, n):
        return ""
Fore = Style = Back = _C()
def init(*a, **k):
    pass
