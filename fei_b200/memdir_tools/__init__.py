"""Reference-shaped API (memdir_tools.search / filter / memorychain) over libfeiscan."""
