"""Command line shim with the reference's argv words (reference ``traffic_classifier.py:174-246``).

    python -m traffic_classifier_sdn_b200.cli <subcommand> [options]

Subcommands: the reference's ``train <TypeOfData>``, ``logistic``, ``kmeans``, ``knearest``, ``svm``,
``Randomforest``, ``gaussiannb`` -- plus ``kneighbors`` (the spelling the reference's own loader branch
tests, ``:235``; with ``knearest`` the reference crashes on an unbound ``infile``) and the README's
``supervised`` / ``unsupervised`` (-> ``logistic`` / ``kmeans``, the only split the reference draws,
``:106-108``).  The model is read from ``models/<Name>`` exactly like ``:229-244`` but as data only, packed to
HBM once, and every report classifies ALL flows with ONE ``model.predict`` on the GPU (the reference
calls ``predict`` once per flow, ``:103-106``).

Options (none exist in the reference; defaults reproduce it):
    --monitor-cmd CMD   child process whose stdout carries the monitor lines
                        (default ``sudo ryu run simple_monitor_13.py``, reference ``:22``; use e.g.
                        ``cat capture.log`` to replay a recorded monitor log -- Ryu is not needed then)
    --models DIR        directory holding the six model files (default ``models``)
    --every N           classify every N-th input line (default 10, reference ``:167``)
    --timeout S         training capture length in seconds (default 900, reference ``:27``)
    --host-features     keep the flow table and its feature derivation on the host (default: the table's state lives in
                        HBM, ``Flow.updateforward/updatereverse`` run in ``tcsdn_flow_update`` and ``predict`` reads the
                        features where that kernel wrote them)
"""
from __future__ import annotations

import os
import signal
import subprocess
import sys

from . import flows as _flows

CMD = "sudo ryu run simple_monitor_13.py"
TIMEOUT = 15 * 60
SUBCOMMANDS = ("train", "logistic", "kmeans", "knearest", "svm", "Randomforest", "gaussiannb")
ALIASES = {"knearest": "kneighbors", "supervised": "logistic", "unsupervised": "kmeans"}
# cluster id / class index -> name, as hard-coded at reference :109-114
INT_LABELS = {0: "dns", 1: "game", 2: "ping", 3: "quake", 4: "telnet", 5: "voice"}
FIELDS = ["Flow ID", "Src MAC", "Dest MAC", "Traffic Type", "Forward Status", "Reverse Status"]


def print_help(out=None):
    w = (out or sys.stdout).write
    w("\nUsage: sudo python traffic_classifier.py [subcommand] [options]\n")
    w("\n\tTo collect training data for a certain type of traffic, run: sudo python traffic_classifier.py train <TypeOfData>\n")
    w("\n\tTo start a near real time traffic classification application using unsupervised ML, run: sudo python traffic_classifier.py <NameOfAlgo>\n")
    w("\n\tTo start a near real time traffic classification application using supervised ML, run: sudo python traffic_classifier.py <NameOfAlgo>\n")
    w("\n\t Available algorithms Logistic Regression, K Means clustering, K nearest neighbors, Random Forest Classifier, SVM, Gaussian Naive Bayes\n")
    w("\n\t SUBCOMMANDS = ('train', 'logistic', 'kmeans', 'knearest', 'svm', 'Randomforest', 'gaussiannb')\n")


def render_table(rows, out=None):
    """ASCII table in PrettyTable's default style (prettytable is not a dependency here)."""
    out = out or sys.stdout
    cells = [FIELDS] + [[str(c) for c in r] for r in rows]
    widths = [max(len(r[i]) for r in cells) for i in range(len(FIELDS))]
    sep = "+" + "+".join("-" * (w + 2) for w in widths) + "+"
    lines = [sep, "|" + "|".join(" " + c.center(w) + " " for c, w in zip(cells[0], widths)) + "|", sep]
    for r in cells[1:]:
        lines.append("|" + "|".join(" " + c.center(w) + " " for c, w in zip(r, widths)) + "|")
    lines.append(sep)
    out.write("\n".join(lines) + "\n")


def classify_table(table: _flows.FlowTable, model):
    """reference printclassifier (:99-118): one row per flow with the predicted traffic type."""
    if len(table) == 0:
        return []
    if hasattr(table, "features_device"):      # update -> predict on the GPU: the features never leave HBM
        idx = model.predict_indices(table.features_device())
        labels = model._labels_from_indices(idx.cpu().numpy())
        model.sync_check()
    else:
        labels = model.predict(table.features())   # ONE batched call instead of one per flow
    rows = []
    for (fid, src, dst, fwd, rev), lab in zip(table.rows(), labels):
        name = lab
        try:  # integer outputs (KMeans cluster ids) go through the reference's fixed map :109-114
            if not isinstance(lab, str) and int(lab) == lab and int(lab) in INT_LABELS:
                name = INT_LABELS[int(lab)]
        except (TypeError, ValueError):
            pass
        rows.append((fid, src, dst, name, fwd, rev))
    return rows


def run_monitor(stream, model=None, traffic_type=None, f=None, every=10, out=None, max_lines=None, device_table=False):
    """reference run_ryu (:144-171) over any binary line stream.  Unlike the reference, which never leaves its
    loop (``out == ''`` compares bytes to str, :150), this returns at end of stream.
    device_table: keep the flow state in HBM and derive the features there (flows.DeviceFlowTable)."""
    table = _flows.DeviceFlowTable() if device_table else _flows.FlowTable()
    count = 0
    while True:
        line = stream.readline()
        if not line:
            break
        rec = _flows.parse_monitor_line(line)
        if rec is not None:
            table.ingest(rec)
            if model is not None:
                if count % every == 0:
                    render_table(classify_table(table, model), out)
            elif f is not None:
                for ln in table.training_lines(traffic_type):
                    f.write(ln)
        count += 1
        if max_lines is not None and count >= max_lines:
            break
    return table


def _alarm_handler(signum, frame):
    print("Finished collecting data.")
    raise TimeoutError()


def _pop_option(argv, name, default):
    if name in argv:
        i = argv.index(name)
        if i + 1 >= len(argv):
            raise SystemExit(f"ERROR: {name} needs a value")
        val = argv[i + 1]
        del argv[i:i + 2]
        return val
    return default


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    monitor_cmd = _pop_option(argv, "--monitor-cmd", CMD)
    models_dir = _pop_option(argv, "--models", "models")
    every = int(_pop_option(argv, "--every", 10))
    timeout = int(_pop_option(argv, "--timeout", TIMEOUT))
    host_features = "--host-features" in argv
    if host_features:
        argv.remove("--host-features")
    if len(argv) < 1:
        print("ERROR: Incorrect # of args")
        print()
        print_help()
        return 0
    word = argv[0]
    known = SUBCOMMANDS + tuple(ALIASES) + ("kneighbors",)
    if len(argv) == 1 and word not in known:
        print("ERROR: Unknown subcommand argument.")
        print("       Currently subaccepted commands are: %s" % str(SUBCOMMANDS).strip("()"))
        print()
        print_help()
        return 0
    if word == "train":
        if len(argv) != 2:
            print("ERROR: specify traffic type.\n")
            return 0
        traffic_type = argv[1]
        p = subprocess.Popen(monitor_cmd, shell=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                             start_new_session=True)
        with open(traffic_type + "_training_data.csv", "w") as f:
            signal.signal(signal.SIGALRM, _alarm_handler)
            signal.alarm(timeout)
            try:
                f.write(_flows.TRAINING_HEADER)
                run_monitor(p.stdout, traffic_type=traffic_type, f=f)
            except TimeoutError:
                print("Exiting")
            finally:
                signal.alarm(0)
                try:
                    os.killpg(os.getpgid(p.pid), signal.SIGTERM)
                except ProcessLookupError:
                    pass
        return 0
    from . import estimators, modelio
    word = ALIASES.get(word, word)
    if word not in modelio.MODEL_FILES:
        print("ERROR: Unknown subcommand argument.")
        print_help()
        return 0
    model = estimators.load_model(os.path.join(models_dir, modelio.MODEL_FILES[word]))
    p = subprocess.Popen(monitor_cmd, shell=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         start_new_session=True)
    try:
        run_monitor(p.stdout, model=model, every=every, device_table=not host_features)
    finally:
        try:
            os.killpg(os.getpgid(p.pid), signal.SIGTERM)
        except ProcessLookupError:
            pass
    return 0


if __name__ == "__main__":
    sys.exit(main())
