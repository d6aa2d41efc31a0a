"""extinction.fitzpatrick99 stand-in: the oracle's restatement of the published curve (parity unpinned)."""
from oracle.reference_port import fitzpatrick99  # noqa: F401
