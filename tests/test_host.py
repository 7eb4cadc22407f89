"""Host-side mirror of the reference interface: FASTA wire format (FASTAFile.cpp:71-187),
residue-cell filter (Sequence.cpp:57-69), AlignmentTool surface (AlignmentTool.h:36-71) and the
settings element (ClustalO.cpp:54-86)."""
import xml.etree.ElementTree as ET

import numpy as np

from tweakseq_b200 import AlignmentTool, B200Gotoh, synth
from tweakseq_b200.backend import square
from tweakseq_b200.fasta import (EXCLUDE_CELL, HIGHLIGHT_CELL, filter_cells, parse_comment, read_fasta,
                                 write_fasta)


def test_filter_cells_follows_sequence_filter():
    cells = [ord("A"), ord("C") | EXCLUDE_CELL, ord("-"), ord("D") | HIGHLIGHT_CELL, ord("E") | EXCLUDE_CELL | HIGHLIGHT_CELL]
    assert filter_cells(cells, True) == "A-D"          # excluded dropped, flags stripped, '-' kept
    assert filter_cells(cells, False) == "AC-DE"


def test_parse_comment_label_rule():
    assert parse_comment(">sp|P1|X some protein") == "sp|P1|X"
    assert parse_comment(">abc") == "abc"
    assert parse_comment(";legacy header") == "legacy"


def test_fasta_round_trip_and_reader_state_machine(tmp_path):
    p = tmp_path / "x.fa"
    labels = ["s1", "s2", "s3"]
    seqs = ["A" * 200, "ACDEFG", "W" * 81]
    write_fasta(str(p), labels, seqs, [">s1 first", ">s2", ">s3 third one"])
    lines = p.read_text().splitlines()
    assert lines[0] == ">s1 first" and len(lines[1]) == 80 and len(lines[3]) == 40   # 80-column residue lines
    l2, s2, c2 = read_fasta(str(p))
    assert l2 == labels and s2 == seqs and c2[2] == ">s3 third one"
    q = tmp_path / "y.fa"
    q.write_text("junk before any header\n>a desc\n;extra comment\n\n  ACD  \nEFG\n;b old style\nKLM\n")
    l3, s3, _ = read_fasta(str(q))
    assert l3 == ["a", "b"] and s3 == ["ACDEFG", "KLM"]


def test_alignment_tool_defaults_and_noop_virtuals():
    t = AlignmentTool()
    assert t.name() == "" and t.version() == "" and t.executable() == ""
    assert t.preferred() is False and t.usesStdOut() is False and t.inProcess() is False
    t.setExecutable("/x/y"); t.setPreferred(True)
    assert t.executable() == "/x/y" and t.preferred()
    assert t.makeCommand("in", "out") == ("", [])


def test_b200gotoh_settings_round_trip_like_clustalo():
    t = B200Gotoh()
    assert t.name() == "b200gotoh" and t.inProcess() and not t.usesStdOut()
    assert t.executable().endswith("libtsqb200.so")
    t.setPreferred(True); t.gap_open, t.gap_extend, t.device = 9, 2, 1
    assert t.align
    t.align = False
    root = ET.Element("settings")
    other = ET.SubElement(root, "alignment_tool")       # another tool's element must be ignored
    ET.SubElement(other, "name").text = "clustalo"
    ET.SubElement(other, "path").text = "/usr/local/bin/clustalo"
    ET.SubElement(other, "preferred").text = "no"
    t.writeSettings(root)
    e = root.findall("alignment_tool")[1]
    assert [c.tag for c in e][:3] == ["name", "path", "preferred"] and e.find("preferred").text == "yes"
    u = B200Gotoh()
    u.readSettings(root)
    assert u.preferred() and (u.gap_open, u.gap_extend, u.device) == (9, 2, 1) and not u.align
    assert u.executable() == t.executable() and "tsq-b200" in u.version()


def test_synthetic_configs_are_seeded_and_shaped():
    a1, s1 = synth.config(2, 0.02)
    a2, s2 = synth.config(2, 0.02)
    assert s1 == s2 and a1 == 0 and len(s1) == 20 and all(len(x) == 300 for x in s1)
    _, c1 = synth.config(1)
    assert len(c1) == 100 and all(200 <= len(x) <= 400 for x in c1)
    a4, c4 = synth.config(4, 0.01)
    assert a4 == 1 and all(10000 <= len(x) <= 30000 and set(x) <= set("ACGT") for x in c4)
    assert synth.total_cells(["AAA", "CC", "G"]) == 3 * 2 + 3 * 1 + 2 * 1


def test_square_from_packed():
    m = square(np.array([1, 2, 3], dtype=np.int32), 3)
    assert m.tolist() == [[0, 1, 2], [1, 0, 3], [2, 3, 0]]
