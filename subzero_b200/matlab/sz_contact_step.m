function [Floe, kill, transfer, ghosts] = sz_contact_step(Floe, floebound, c2_boundary, dt, Nb, COLLISION, PERIODIC, Modulus)
%SZ_CONTACT_STEP  GPU replacement of floe_interactions_all.m lines 16-277 (ghost floes, potential interactions,
% floe_interactions over every pair and the walls, mirror, torques, force/torque sums, periodic wrap).
%
% Called by this folder's floe_interactions_all.m right after the dead floes have been dropped (:12-14).  On return every
% floe i > Nb carries the fields the reference writes:
%   interactions (K x 7: [partner Fx Fy Px Py torque overlap]), OverlapArea, collision_force (1 x 2),
%   collision_torque, Stress = zeros(2), potentialInteractions = [], alive, Xi, Yi (wrapped),
% and kill / transfer (1 x N0) are what :138-145,175-179 compute.  Ghost floes are not appended to Floe: their forces are
% already folded into their parents (:242-245).  `ghosts` describes them for a caller that needs the structs (the ridging /
% rafting tail): parent (1-based index in the extended list), x, y (shifted centroid), interactions (cell of K x 7),
% overlap_area, fx, fy, torque.
    N0 = numel(Floe);
    prm = struct('Lx', max(c2_boundary(1,:)), 'Ly', max(c2_boundary(2,:)), 'modulus', Modulus, 'dt', dt, ...
                 'Nb', Nb, 'periodic', double(PERIODIC), 'collision', double(COLLISION));
    nv = arrayfun(@(f) size(f.c_alpha, 2), Floe);
    ca = [Floe.c_alpha];                                    % 2 x V, closed outlines back to back
    soa = struct('x', cat(1, Floe.Xi), 'y', cat(1, Floe.Yi), 'rmax', cat(1, Floe.rmax), 'h', cat(1, Floe.h), ...
                 'area', cat(1, Floe.area), 'u', cat(1, Floe.Ui), 'v', cat(1, Floe.Vi), 'ksi', cat(1, Floe.ksi_ice), ...
                 'alive', double(cat(1, Floe.alive)), 'voff', [0; cumsum(nv(:))], 'vx', ca(1,:)', 'vy', ca(2,:)');
    if PERIODIC
        out = sz_contact_mex(prm, soa);
    else
        hv = holes(floebound.poly).Vertices;                 % floe_interactions.m:31
        bnd = struct('x', hv(:,1), 'y', hv(:,2), 'box_x', c2_boundary(1,:)', 'box_y', c2_boundary(2,:)', ...
                     'area', floebound.area, 'h', floebound.h, 'xi', floebound.Xi, 'yi', floebound.Yi, ...
                     'u', floebound.Ui, 'v', floebound.Vi, 'ksi', floebound.ksi_ice);   % floe_interactions.m:171 reads them (zeros in Subzero.m)
        out = sz_contact_mex(prm, soa, bnd);
    end
    for i = 1+Nb:N0
        r = out.row_off(i)+1 : out.row_off(i+1);
        Floe(i).interactions = out.rows(:, r)';
        Floe(i).OverlapArea = out.overlap_area(i);
        Floe(i).collision_force = [out.fx(i) out.fy(i)];
        Floe(i).collision_torque = out.torque(i);
        Floe(i).Stress = zeros(2);
        Floe(i).potentialInteractions = [];
        Floe(i).alive = out.alive(i);
        Floe(i).Xi = out.xi(i);  Floe(i).Yi = out.yi(i);
    end
    kill = out.kill';  transfer = out.transfer';
    ng = numel(out.ghost_parent);
    ghosts = struct('parent', out.ghost_parent, 'x', out.ghost_x, 'y', out.ghost_y, 'fx', out.ghost_fx, 'fy', out.ghost_fy, ...
                    'torque', out.ghost_torque, 'overlap_area', out.ghost_overlap_area);
    ghosts.interactions = cell(ng, 1);
    for k = 1:ng
        r = out.row_off(N0+k)+1 : out.row_off(N0+k+1);
        ghosts.interactions{k} = out.rows(:, r)';
    end
end
