"""Native host runtime (C++): batch assembler for pinned-host datasets."""
