"""Pick the checker: the compiled reference (oracle/_ref) when it is present, otherwise the CPU restatement
(oracle/port).  Both expose the same functions; tests/test_oracle.py pins them against each other and against the
golden vectors."""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def get_oracle(prefer="ref"):
    from oracle import port, ref
    if prefer == "ref" and ref.available():
        try:
            ref.lib()
            return ref
        except OSError:
            pass
    port.lib()
    return port


def both():
    from oracle import port, ref
    out = [port]
    if ref.available():
        out.append(ref)
    return out
