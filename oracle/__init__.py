"""TEST INFRASTRUCTURE ONLY — see oracle/README.md.  Never imported by zpack_b200/."""
