"""The socket / client compatibility layer on CPU: the wire protocol of
/root/reference/delivery_drone/game/socket_server.py:126-263 and socket_client.py:31-224, served
from duck-typed game objects (here the oracle's reference-faithful python port, so no GPU)."""
import importlib
import json
import socket

import pytest

from oracle.drone_port import PortDroneGame

dd = importlib.import_module("reinforcement-learning-101_b200")
compat = importlib.import_module("reinforcement-learning-101_b200.compat")


class _Game(PortDroneGame):
    """PortDroneGame with the reference's `_get_info` name."""

    def _get_info(self):
        return self.info()


@pytest.fixture()
def server():
    games = [_Game(None, False, False) for _ in range(3)]
    srv = compat.DroneSocketServer(games, host="127.0.0.1", port=0)
    srv.start(background=True)
    yield srv, games
    srv.stop()


def test_client_server_roundtrip_matches_direct_calls(server):
    srv, games = server
    twin = _Game(None, False, False)
    with compat.DroneGameClient("127.0.0.1", srv.port, timeout=10) as c:
        assert c.num_games == 3
        s0 = c.reset(1)
        t0 = twin.reset()
        assert isinstance(s0, compat.DroneState) and s0.drone_x == t0["drone_x"] and s0.steps == 0
        total = 0.0
        for k in range(400):
            a = {"main_thrust": int(k % 3 == 0), "right_thrust": int(k % 7 == 0)}      # left_thrust missing -> 0
            st, r, done, info = c.step(a, game_id=1)
            ts, tr, td, ti = twin.step(a)
            assert st == compat.DroneState(**ts) and r == float(tr) and done == bool(td)
            assert info == json.loads(json.dumps(ti))
            total += r
            if done:
                break
        assert done and sorted(info) == sorted(compat.INFO_KEYS)
        # step after done: reward 0, done True, full info + needs_reset (BUGFIX.md:40-52)
        st2, r2, d2, i2 = c.step({"main_thrust": 1}, game_id=1)
        assert (r2, d2, i2["needs_reset"], st2.steps) == (0, True, True, st.steps)
        assert sorted(k for k in i2 if k != "needs_reset") == sorted(compat.INFO_KEYS)
        # other games untouched; GET_STATE works and does not step
        g0 = c.get_state(0)
        assert g0.steps == 0 and games[0].steps == 0 and games[2].steps == 0
        assert c.get_state(1) == st2
        # RESET starts a new episode
        assert c.reset(1).steps == 0 and games[1].episode == 2


def test_client_errors(server):
    srv, _ = server
    c = compat.DroneGameClient("127.0.0.1", srv.port, timeout=10)
    with pytest.raises(RuntimeError, match="Not connected"):
        c.step({}, 0)
    with pytest.raises(RuntimeError, match="Not connected"):
        c.get_state(0)
    c.reset(0)                                   # lazy connect
    with pytest.raises(ValueError, match="Invalid game_id"):
        c.reset(3)
    with pytest.raises(ValueError):
        c.step({}, -1)
    c.close()
    assert not c.connected


def test_raw_wire_format(server):
    srv, _ = server
    s = socket.create_connection(("127.0.0.1", srv.port), timeout=10)
    f = s.makefile("rwb")

    def ask(obj=None, raw=None):
        f.write(raw if raw is not None else (json.dumps(obj) + "\n").encode())
        f.flush()
        return json.loads(f.readline())

    assert json.loads(f.readline()) == {"type": "HANDSHAKE", "num_games": 3}
    r = ask({"type": "RESET"})                   # game_id defaults to 0
    assert r["type"] == "STATE" and r["game_id"] == 0 and r["reward"] == 0.0 and r["done"] is False and r["info"] == {}
    assert list(r["state"]) == list(compat.STATE_KEYS)
    r = ask({"type": "STEP", "game_id": 2, "action": {"main_thrust": 1}})
    assert r["game_id"] == 2 and r["state"]["steps"] == 1 and r["state"]["drone_fuel"] == 0.998
    assert list(r["info"]) == list(compat.INFO_KEYS)
    r = ask({"type": "GET_STATE", "game_id": 2})
    assert r["reward"] == 0.0 and r["done"] is False and r["info"]["steps"] == 1
    assert ask({"type": "STEP", "game_id": 9})["type"] == "ERROR"
    assert "Unknown message type" in ask({"type": "DANCE"})["message"]
    assert "Invalid JSON" in ask(raw=b"{nope\n")["message"]
    f.write(b'{"type": "CLOSE"}\n'); f.flush()
    assert f.readline() == b""                   # server ends the session
    s.close()
    # the server accepts the next client
    with compat.DroneGameClient("127.0.0.1", srv.port, timeout=10) as c:
        assert c.num_games == 3


def test_in_process_client_and_aliases():
    games = [_Game(None, False, False) for _ in range(2)]
    c = compat.InProcessDroneGameClient(games)
    with pytest.raises(RuntimeError):
        c.step({}, 0)
    s = c.reset(0)
    assert c.connected and s.drone_y == 100 / 600
    st, r, d, info = c.step({"left_thrust": "yes"}, 0)        # any truthy value (game_engine.py:114-118)
    assert games[0].angvel == pytest.approx(-0.3 * 0.95) and st.steps == 1
    with pytest.raises(ValueError):
        c.get_state(2)
    with pytest.raises(TypeError):
        compat.DroneState(**{**games[0].get_state(), "extra": 1})
    compat.install_aliases()
    from delivery_drone.game.socket_client import DroneGameClient, DroneState
    from delivery_drone.game.socket_server import GameSocketServer
    assert DroneGameClient is compat.DroneGameClient and DroneState is compat.DroneState
    assert GameSocketServer is compat.DroneSocketServer
    assert compat.action_bits({"main_thrust": 1, "right_thrust": True}) == 5 and compat.action_bits(None) == 0
