"""Writes a small netconvert-style SUMO scenario (two signalised 4-way junctions in a row, one lane per direction):
`mini.net.xml` (edges, internal lanes, junction right-of-way matrices, connections with tl / linkIndex / via, tlLogic
programs with yellow phases) and `mini.rou.xml` (flows given by from/to/via + vehsPerHour, by period, by number and by a
named route).  Test input for the general ingest (net/sumo_ingest.py); not derived from the reference's files."""
import os

L_EDGE, L_INT = 120.0, 9.0
J = {"A": ("W", "B", "NA", "SA"), "B": ("A", "E", "NB", "SB")}      # junction: (west, east, north, south) neighbours


def write(dirname):
    edges, internal, conns, juncs, tls = [], [], [], [], []
    for j, (w, e, n, s) in J.items():
        for o in (w, e, n, s):
            edges.append((o, j)); edges.append((j, o))
        # approaches clockwise from north; per approach right, straight, left   (link index = 3 * approach + turn)
        appr = [(n, (w, s, e)), (e, (n, w, s)), (s, (e, n, w)), (w, (s, e, n))]
        ints, reqs = [], []
        for a, (frm, tos) in enumerate(appr):
            for t, to in enumerate(tos):
                li = 3 * a + t
                via = ":%s_%d_0" % (j, li)
                internal.append((via, L_INT))
                conns.append(dict(frm="%s_%s" % (frm, j), to="%s_%s" % (j, to), via=via, tl=j, li=li,
                                  dir="rsl"[t], state="o"))
                ints.append(via)
        for a in range(4):
            for t in range(3):
                resp = ["0"] * 12
                if t == 2:                                   # left turn yields to the opposing straight and right
                    opp = (a + 2) % 4
                    resp[3 * opp + 1] = "1"; resp[3 * opp + 0] = "1"
                reqs.append("".join(reversed(resp)))
        juncs.append((j, ints, reqs))
        tls.append((j, [("GGgrrrGGgrrr", 30), ("yyyrrryyyrrr", 3), ("rrrGGgrrrGGg", 30), ("rrryyyrrryyy", 3),
                        ("rrrrrrrrrrrr", 2)]))
    seen = set()
    with open(os.path.join(dirname, "mini.net.xml"), "w") as f:
        f.write('<net version="1.1">\n')
        for via, ln in internal:
            f.write('  <edge id="%s" function="internal"><lane id="%s" index="0" speed="8.0" length="%.2f"/></edge>\n'
                    % (via.rsplit("_", 1)[0], via, ln))
        for a, b in edges:
            if (a, b) in seen:
                continue
            seen.add((a, b))
            f.write('  <edge id="%s_%s" from="%s" to="%s"><lane id="%s_%s_0" index="0" speed="13.89" length="%.2f"/></edge>\n'
                    % (a, b, a, b, a, b, L_EDGE))
        for j, prog in tls:
            f.write('  <tlLogic id="%s" type="static" programID="0" offset="0">\n' % j)
            for st, dur in prog:
                f.write('    <phase duration="%d" state="%s"/>\n' % (dur, st))
            f.write('  </tlLogic>\n')
        for j, ints, reqs in juncs:
            f.write('  <junction id="%s" type="traffic_light" intLanes="%s">\n' % (j, " ".join(ints)))
            for i, r in enumerate(reqs):
                f.write('    <request index="%d" response="%s" foes="%s"/>\n' % (i, r, r))
            f.write('  </junction>\n')
        for o in ("W", "E", "NA", "SA", "NB", "SB"):
            f.write('  <junction id="%s" type="dead_end" intLanes=""/>\n' % o)
        for c in conns:
            f.write('  <connection from="%s" to="%s" fromLane="0" toLane="0" via="%s" tl="%s" linkIndex="%d" dir="%s" state="%s"/>\n'
                    % (c["frm"], c["to"], c["via"], c["tl"], c["li"], c["dir"], c["state"]))
            f.write('  <connection from="%s" to="%s" fromLane="0" toLane="0" dir="s" state="M"/>\n'
                    % (c["via"].rsplit("_", 1)[0], c["to"]))
        f.write('</net>\n')
    with open(os.path.join(dirname, "mini.rou.xml"), "w") as f:
        f.write('<routes>\n  <vType id="car" length="5" accel="5" decel="10"/>\n')
        f.write('  <route id="r_ns_b" edges="NB_B B_SB"/>\n')
        f.write('  <flow id="f0" from="W_A" to="B_E" via="A_B" begin="0" end="600" vehsPerHour="600" type="car"/>\n')
        f.write('  <flow id="f1" from="E_B" to="A_W" begin="0" end="600" vehsPerHour="450" type="car"/>\n')
        f.write('  <flow id="f2" from="NA_A" to="A_SA" begin="100" end="500" period="8" type="car"/>\n')
        f.write('  <flow id="f3" route="r_ns_b" begin="0" end="300" number="50" type="car"/>\n')
        f.write('  <flow id="f4" from="NA_A" to="B_SB" begin="0" end="600" vehsPerHour="240" type="car"/>\n')
        f.write('</routes>\n')
    return os.path.join(dirname, "mini.net.xml"), os.path.join(dirname, "mini.rou.xml")


if __name__ == "__main__":
    print(write(os.path.dirname(os.path.abspath(__file__))))
